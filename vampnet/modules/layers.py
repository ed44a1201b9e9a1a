"""vampnet.modules.layers (reference vampnet/modules/layers.py): the names other modules and scripts import."""
from vampnet_b200.modules.transformer import CodebookEmbedding  # noqa: F401  (layers.py:105-164)


def num_params(model):
    """layers.py:31-32."""
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
