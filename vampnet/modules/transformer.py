"""vampnet.modules.transformer (reference vampnet/modules/transformer.py) -> vampnet_b200.modules.transformer.
Only the names callers outside the module use are exported: the model class and its LoRA rank constant."""
from vampnet_b200.modules.transformer import VampNet, LORA_R, CodebookEmbedding  # noqa: F401
