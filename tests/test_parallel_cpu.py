"""CPU, world_size 2, gloo: the N>1 host logic (weight broadcast from rank 0, clip sharding, max-over-ranks)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vampnet_b200 import parallel
    from vampnet_b200.modules.transformer import VampNet
    torch.manual_seed(100 + rank)  # different init per rank
    m = VampNet(n_heads=2, n_layers=1, n_codebooks=4, embedding_dim=128)
    before = torch.cat([p.flatten() for p in m.parameters()]).clone()
    nbytes = parallel.broadcast_module_weights([m], src=0)
    after = torch.cat([p.flatten() for p in m.parameters()])
    gathered = [torch.empty_like(after) for _ in range(world)]
    dist.all_gather(gathered, after)
    # codec: derived packs must be dropped by the broadcast (a stale pack would keep serving rank-local weights)
    from vampnet_b200.codec import DAC
    codec = DAC(encoder_dim=16, decoder_dim=128, n_codebooks=2, precision="fp32")
    codec._pack = {"stale": True}
    with torch.no_grad():
        codec.params.get("encoder.conv1.weight").fill_(float(rank + 1))
    cbytes = parallel.broadcast_module_weights([codec], src=0, bucket_bytes=1 << 16)  # forces several buckets
    codec_ok = codec._pack is None and float(codec.params.get("encoder.conv1.weight").flatten()[0]) == 1.0
    want_bytes = sum(p.numel() * p.element_size() for p in m.parameters())
    lo, hi = parallel.shard_range(7, rank, world)
    mx = parallel.max_over_ranks(float(rank + 1))
    out[rank] = dict(same=all(torch.equal(g, gathered[0]) for g in gathered), changed=not torch.equal(before, after),
                     nbytes=nbytes, shard=(lo, hi), mx=mx, codec_ok=codec_ok, cbytes=cbytes,
                     want_bytes=want_bytes)
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0]["same"] and out[1]["same"]
    assert not out[0]["changed"] and out[1]["changed"]  # rank 1 received rank 0's weights
    assert out[0]["nbytes"] == out[1]["nbytes"] == out[0]["want_bytes"] > 0  # native dtype, every byte once
    assert out[0]["codec_ok"] and out[1]["codec_ok"] and out[0]["cbytes"] == out[1]["cbytes"] > 0
    assert out[0]["shard"] == (0, 4) and out[1]["shard"] == (4, 7)
    assert out[0]["mx"] == out[1]["mx"] == 2.0


def test_shard_range_covers_everything():
    from vampnet_b200.parallel import shard_range
    for n in (0, 1, 8, 31, 256):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
