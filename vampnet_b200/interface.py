"""Interface — drop-in for the reference's ``vampnet.interface.Interface`` surface
(reference vampnet/interface.py:54-562): checkpoint loading, encode, build_mask, chunked coarse_vamp,
coarse_to_fine, vamp, decode.  Same method names, arguments, defaults, return types and error behaviour;
the compute underneath is the sm_100a CUDA path (VampNet.generate, the codec kernels).

Chunk loops are kept (they define the reference's results: every chunk is an independent generate() call
with its own whole-batch N0), but each chunk's loop body is one CUDA-graph replay.
"""
from __future__ import annotations

import logging
import math
from pathlib import Path

import numpy as np
import torch

from . import mask as pmask
from .audio import AudioSignal
from .mask import *  # noqa: F401,F403  (reference does `from .mask import *`, interface.py:13)
from .modules.transformer import VampNet


def signal_concat(audio_signals: list):
    """interface.py:19-24."""
    audio_data = torch.cat([x.audio_data for x in audio_signals], dim=-1)
    return AudioSignal(audio_data, sample_rate=audio_signals[0].sample_rate)


def _load_model(ckpt: str, lora_ckpt: str = None, device: str = "cpu", chunk_size_s: int = 10):
    """interface.py:27-50.  A missing LoRA checkpoint raises instead of blocking on input()."""
    model = VampNet.load(location=Path(ckpt), map_location="cpu", strict=False)
    if lora_ckpt is not None:
        if not Path(lora_ckpt).exists():
            raise FileNotFoundError(f"lora checkpoint {lora_ckpt} does not exist")
        model.load_state_dict(torch.load(lora_ckpt, map_location="cpu"), strict=False)
    model.to(device)
    model.eval()
    model.chunk_size_s = chunk_size_s
    return model


class Interface(torch.nn.Module):
    def __init__(
        self,
        coarse_ckpt: str = None,
        coarse_lora_ckpt: str = None,
        coarse2fine_ckpt: str = None,
        coarse2fine_lora_ckpt: str = None,
        codec_ckpt: str = None,
        wavebeat_ckpt: str = "./models/vampnet/wavebeat.pth",
        device: str = "cpu",
        coarse_chunk_size_s: int = 10,
        coarse2fine_chunk_size_s: int = 3,
        compile=True,
    ):
        super().__init__()
        from .codec import DAC
        assert codec_ckpt is not None, "must provide a codec checkpoint"
        self.codec = DAC.load(Path(codec_ckpt))
        self.codec.eval()
        self.codec.to(device)
        self.codec_path = Path(codec_ckpt)

        assert coarse_ckpt is not None, "must provide a coarse checkpoint"
        self.coarse = _load_model(ckpt=coarse_ckpt, lora_ckpt=coarse_lora_ckpt, device=device,
                                  chunk_size_s=coarse_chunk_size_s)
        self.coarse_path = Path(coarse_ckpt)
        if coarse2fine_ckpt is not None:
            self.c2f_path = Path(coarse2fine_ckpt)
            self.c2f = _load_model(ckpt=coarse2fine_ckpt, lora_ckpt=coarse2fine_lora_ckpt, device=device,
                                   chunk_size_s=coarse2fine_chunk_size_s)
        else:
            self.c2f_path = None
            self.c2f = None
        # WaveBeat (interface.py:96-101) is a separate model outside the hot path
        self.beat_tracker = None
        if wavebeat_ckpt is not None and Path(wavebeat_ckpt).exists():
            logging.debug("wavebeat checkpoint present but the beat tracker is out of scope; beat masks disabled")
        self.device = device
        self.loudness = -24.0
        # `compile` (torch.compile in the reference, interface.py:107-112) is accepted and ignored:
        # there is no tracing compiler on this path, the generate loop is a captured CUDA graph.

    @classmethod
    def from_models(cls, codec, coarse, c2f=None, device="cuda", coarse_chunk_size_s=10, coarse2fine_chunk_size_s=3):
        """Build an Interface around already-constructed modules (tests, benchmarks, synthetic weights)."""
        self = cls.__new__(cls)
        torch.nn.Module.__init__(self)
        self.codec, self.coarse, self.c2f = codec, coarse, c2f
        self.coarse.chunk_size_s = coarse_chunk_size_s
        if c2f is not None:
            self.c2f.chunk_size_s = coarse2fine_chunk_size_s
        self.codec_path = self.coarse_path = self.c2f_path = None
        self.beat_tracker = None
        self.device = device
        self.loudness = -24.0
        return self.to(device)

    # ------------------------------------------------------------------ checkpoints (interface.py:115-174)
    # The reference keeps a local cache under <repo>/models/vampnet ({codec,coarse,c2f}.pth, loras/<name>/{coarse,
    # c2f}.pth) and fills it from the HF hub on a miss (vampnet/__init__.py:19-76).  Here the cache is the only
    # source: $VAMPNET_MODELS_DIR or ./models/vampnet; a miss raises (no network access in this build).
    @staticmethod
    def models_dir() -> Path:
        import os
        return Path(os.environ.get("VAMPNET_MODELS_DIR", "./models/vampnet"))

    @classmethod
    def _cached(cls, *parts: str) -> Path:
        path = cls.models_dir().joinpath(*parts)
        if not path.exists():
            raise RuntimeError(f"{path} is not in the local model cache and this build cannot download it from the HF "
                               f"hub (vampnet/__init__.py:19-59); place the checkpoint there or set VAMPNET_MODELS_DIR")
        return path

    @classmethod
    def default(cls, **kwargs):
        """interface.py:115-126, from the local cache."""
        wavebeat = cls.models_dir() / "wavebeat.pth"
        return cls(coarse_ckpt=cls._cached("coarse.pth"), coarse2fine_ckpt=cls._cached("c2f.pth"),
                   codec_ckpt=cls._cached("codec.pth"), wavebeat_ckpt=str(wavebeat), **kwargs)

    @classmethod
    def available_models(cls):
        """interface.py:128-131: fine-tuned names (those with both coarse.pth and c2f.pth) + "default"."""
        loras = cls.models_dir() / "loras"
        names = sorted(d.name for d in loras.iterdir() if (d / "coarse.pth").exists() and (d / "c2f.pth").exists()) \
            if loras.is_dir() else []
        return names + ["default"]

    def load_finetuned(self, name: str):
        """interface.py:134-144."""
        assert name in self.available_models(), f"{name} is not a valid model name"
        where = () if name == "default" else ("loras", name)
        self.reload(coarse_ckpt=self._cached(*where, "coarse.pth"), c2f_ckpt=self._cached(*where, "c2f.pth"))

    def reload(self, coarse_ckpt: str = None, c2f_ckpt: str = None):
        """Swap checkpoints (interface.py:146-174); a checkpoint already loaded is skipped.  A checkpoint of the same
        architecture is hot-swapped into the live model (VampNet.swap_checkpoint: packed device buffers rewritten in
        place, workspaces / tensor maps / captured generate graphs kept); otherwise the model is rebuilt."""
        for attr, path_attr, ckpt in (("coarse", "coarse_path", coarse_ckpt), ("c2f", "c2f_path", c2f_ckpt)):
            if ckpt is None or getattr(self, path_attr) == Path(ckpt):
                continue
            model = getattr(self, attr)
            if model is None or not model.swap_checkpoint(ckpt):
                chunk_size_s = model.chunk_size_s if model is not None else (10 if attr == "coarse" else 3)
                setattr(self, attr, _load_model(ckpt=ckpt, device=self.device, chunk_size_s=chunk_size_s))
            setattr(self, path_attr, Path(ckpt))

    # ------------------------------------------------------------------ unit helpers (interface.py:176-201)
    def s2t(self, seconds: float):
        """seconds to tokens"""
        if isinstance(seconds, np.ndarray):
            return np.ceil(seconds * self.codec.sample_rate / self.codec.hop_length)
        return math.ceil(seconds * self.codec.sample_rate / self.codec.hop_length)

    def s2t2s(self, seconds: float):
        return self.t2s(self.s2t(seconds))

    def t2s(self, tokens: int):
        return tokens * self.codec.hop_length / self.codec.sample_rate

    def to(self, device):
        self.device = device
        self.coarse.to(device)
        self.codec.to(device)
        if self.c2f is not None:
            self.c2f.to(device)
        return self

    def set_chunk_size(self, chunk_size_s: float):
        self.coarse.chunk_size_s = chunk_size_s

    # ------------------------------------------------------------------ codec boundary (interface.py:203-224)
    def decode(self, z: torch.Tensor):
        return self.coarse.decode(z, self.codec)

    def _preprocess(self, signal: AudioSignal):
        signal = (signal.clone().resample(self.codec.sample_rate).to_mono().normalize(self.loudness)
                  .ensure_max_of_audio(1.0))
        signal.samples, length = self.codec.preprocess(signal.samples, signal.sample_rate)
        return signal

    @torch.inference_mode()
    def encode(self, signal: AudioSignal):
        signal = signal.to(self.device)
        signal = self._preprocess(signal)
        return self.codec.encode(signal.samples, signal.sample_rate)["codes"]

    def make_beat_mask(self, *a, **k):
        raise RuntimeError("make_beat_mask needs the WaveBeat tracker (interface.py:226-322), a separate model "
                           "outside the hot path (SURVEY.md §2 row 9)")

    # ------------------------------------------------------------------ chunk helpers
    @staticmethod
    def _spans(total: int, span: int):
        """[lo, hi) frame ranges of consecutive chunks of `span` frames covering `total` frames."""
        return [(lo, min(lo + span, total)) for lo in range(0, total, span)]

    # ------------------------------------------------------------------ coarse -> fine (interface.py:327-380)
    @torch.inference_mode()
    def coarse_to_fine(self, z: torch.Tensor, mask: torch.Tensor = None, return_mask: bool = False, **kwargs):
        """Fill the fine codebooks given the coarse ones, c2f.chunk_size_s seconds at a time.  The sequence is
        zero-padded to a whole number of chunks (padding frames masked), missing codebooks are appended as zeros and
        the conditioning codebooks are never masked; every chunk is an independent generate() call."""
        assert self.c2f is not None, "No coarse2fine model loaded"
        n_frames = z.shape[-1]
        span = self.s2t(self.c2f.chunk_size_s)
        tail = (-n_frames) % span
        if tail:
            z = torch.nn.functional.pad(z, (0, tail))
            if mask is not None:
                mask = torch.nn.functional.pad(mask, (0, tail), value=1)
        missing = self.c2f.n_codebooks - z.shape[1]
        if missing > 0:
            z = torch.cat([z, z.new_zeros(z.shape[0], missing, z.shape[-1])], dim=1)
        if mask is not None:
            mask = mask.clone()
            mask[:, :self.c2f.n_conditioning_codebooks, :] = 0
        parts = []
        for lo, hi in self._spans(z.shape[-1], span):
            parts.append(self.c2f.generate(codec=self.codec, time_steps=span, start_tokens=z[..., lo:hi],
                                           mask=None if mask is None else mask[..., lo:hi], return_signal=False,
                                           cfg_guidance=None, **kwargs))
        fine = torch.cat(parts, dim=-1)
        result = fine[..., :n_frames].clone()
        if not return_mask:
            return result
        remasked, _ = pmask.apply_mask(fine, mask, self.c2f.mask_token)
        return result, remasked[..., :n_frames].clone()

    # ------------------------------------------------------------------ coarse (interface.py:382-452)
    @torch.inference_mode()
    def coarse_vamp(self, z, mask, return_mask=False, gen_fn=None, **kwargs):
        """Regenerate the masked coarse tokens, coarse.chunk_size_s seconds at a time.  A chunk that keeps at least one
        frame also keeps its first and last frame as anchors so that stitched chunks do not jump
        (interface.py:407-413).  Fine codebooks ride along untouched."""
        n_books = self.coarse.n_codebooks
        tokens, keep = z[:, :n_books, :].clone(), mask[:, :n_books, :]
        assert keep.dtype == torch.long, f"mask must be long dtype, but got {keep.dtype}"
        assert bool(((keep == 0) | (keep == 1)).all()), "mask must be binary"   # the one host sync of this call
        span = self.s2t(self.coarse.chunk_size_s)
        run = gen_fn or self.coarse.generate
        starts, results = [], []
        for lo, hi in self._spans(tokens.shape[-1], span):
            # a chunk that keeps anything also keeps its first and last frame; decided on the device (no sync):
            # edge value = 0 where any(m == 0) else unchanged
            m = keep[..., lo:hi].clone()
            anchors = (m == 0).any()
            m[..., 0] = torch.where(anchors, torch.zeros_like(m[..., 0]), m[..., 0])
            m[..., -1] = torch.where(anchors, torch.zeros_like(m[..., -1]), m[..., -1])
            start, m = pmask.apply_mask(tokens[..., lo:hi], m, self.coarse.mask_token, check=False)
            starts.append(start)
            results.append(run(codec=self.codec, time_steps=span, start_tokens=start, mask=m, return_signal=False,
                               **kwargs))
        out = torch.cat([torch.cat(results, dim=-1), z[:, n_books:, :]], dim=1)
        return (out, torch.cat(starts, dim=-1)) if return_mask else out

    # ------------------------------------------------------------------ masks (interface.py:454-489)
    def build_mask(self, z: torch.Tensor, sig: AudioSignal = None, rand_mask_intensity: float = 1.0,
                   prefix_s: float = 0.0, suffix_s: float = 0.0, periodic_prompt: int = 7,
                   periodic_prompt_width: int = 1, onset_mask_width: int = 0, _dropout: float = 0.0,
                   upper_codebook_mask: int = 3, ncc: int = 0):
        """1 = regenerate, 0 = keep.  Intersection (mask_and) of: Bernoulli(rand_mask_intensity), kept prefix/suffix,
        periodic prompt (random roll), optional onset prompt; then time dropout, `ncc` always-kept codebooks and every
        codebook >= upper_codebook_mask fully masked.  Order and RNG use follow the reference exactly."""
        layers = [
            pmask.linear_random(z, rand_mask_intensity),
            pmask.inpaint(z, self.s2t(prefix_s), self.s2t(suffix_s)),
            pmask.periodic_mask(z, periodic_prompt, periodic_prompt_width, random_roll=True),
        ]
        if onset_mask_width > 0:
            assert sig is not None, "must provide a signal to use onset mask"
            layers.append(pmask.onset_mask(sig, z, self, width=onset_mask_width))
        mask = layers[0]
        for other in layers[1:]:
            mask = pmask.mask_and(mask, other)
        mask = pmask.codebook_unmask(pmask.dropout(mask, _dropout), ncc)
        return pmask.codebook_mask(mask, int(upper_codebook_mask), None)

    # ------------------------------------------------------------------ vamp (interface.py:491-562)
    def vamp(self, codes: torch.Tensor, mask: torch.Tensor, batch_size: int = 1, feedback_steps: int = 1,
             time_stretch_factor: int = 1, return_mask: bool = False, **kwargs):
        """codes, mask (1|B, 14, T) -> (B, 14, T'): coarse stage (`feedback_steps` passes, kwargs forwarded to
        generate) then the fine stage, which the reference pins to 2 sampling steps with default temperature
        (interface.py:545-551).  time_stretch_factor k > 1 inserts k-1 always-masked frames after every frame."""
        z = codes.expand(batch_size, -1, -1)
        mask = mask.expand(batch_size, -1, -1)
        k = int(time_stretch_factor)
        if k > 1:
            z = z.repeat_interleave(k, dim=-1)
            inserted = torch.ones_like(z)
            inserted[..., ::k] = 0
            mask = (mask.repeat_interleave(k, dim=-1).bool() | inserted.bool()).long()
        n_coarse = self.coarse.n_codebooks
        zv = z
        for i in range(feedback_steps):
            zv, coarse_start = self.coarse_vamp(zv, mask=mask, return_mask=True, **kwargs)
            coarse_start = coarse_start.roll(shifts=(i + 1) % feedback_steps, dims=-1)
        if zv.shape[1] < z.shape[1]:
            zv = torch.cat([zv, z[:, n_coarse:, :]], dim=1)
        zv, fine_start = self.coarse_to_fine(zv, mask=mask, typical_filtering=True, _sampling_steps=2, return_mask=True)
        if not return_mask:
            return zv
        return zv, torch.cat([coarse_start[:, :n_coarse, :], fine_start[:, n_coarse:, :]], dim=1).cpu()
