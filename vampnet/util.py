"""vampnet.util (reference vampnet/util.py): the three tensor helpers callers import, as plain torch views."""
import torch


def scalar_to_batch_tensor(x, batch_size):
    """util.py:6-7."""
    return torch.tensor(x).repeat(batch_size)


def codebook_flatten(tokens: torch.Tensor):
    """(batch, codebook, time) -> (batch, time * codebook), s = t*C + c  (util.py:35-39, "b c t -> b (t c)")."""
    return tokens.permute(0, 2, 1).reshape(tokens.shape[0], -1)


def codebook_unflatten(flat_tokens: torch.Tensor, n_c: int = None):
    """(batch, time * codebook) -> (batch, codebook, time)  (util.py:41-46)."""
    return flat_tokens.reshape(flat_tokens.shape[0], -1, n_c).permute(0, 2, 1)


def parallelize(fn, *iterables, parallel: str = "thread_map", **kwargs):
    """util.py:10-33 without the tqdm progress bars."""
    if parallel == "thread_map":
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=kwargs.get("max_workers")) as ex:
            return list(ex.map(fn, *iterables))
    if parallel == "process_map":
        from concurrent.futures import ProcessPoolExecutor
        with ProcessPoolExecutor(max_workers=kwargs.get("max_workers")) as ex:
            return list(ex.map(fn, *iterables))
    if parallel == "single":
        return [fn(*xs) for xs in zip(*iterables)]
    raise ValueError(f"parallel must be one of 'thread_map', 'process_map', 'single', but got {parallel}")
