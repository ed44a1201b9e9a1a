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
    @classmethod
    def default(cls):
        raise RuntimeError("Interface.default() downloads checkpoints from the HF hub (vampnet/__init__.py:19-59); "
                           "this build has no network access. Construct Interface(...) with local checkpoint paths.")

    @classmethod
    def available_models(cls):
        return ["default"]

    def load_finetuned(self, name: str):
        raise RuntimeError("load_finetuned() downloads from the HF hub; use reload(coarse_ckpt, c2f_ckpt) with local files")

    def reload(self, coarse_ckpt: str = None, c2f_ckpt: str = None):
        """Swap checkpoints in place (interface.py:146-174)."""
        if coarse_ckpt is not None and Path(coarse_ckpt) != self.coarse_path:
            self.coarse = _load_model(ckpt=coarse_ckpt, device=self.device, chunk_size_s=self.coarse.chunk_size_s)
            self.coarse_path = Path(coarse_ckpt)
        if c2f_ckpt is not None and Path(c2f_ckpt) != self.c2f_path:
            self.c2f = _load_model(ckpt=c2f_ckpt, device=self.device, chunk_size_s=self.c2f.chunk_size_s)
            self.c2f_path = Path(c2f_ckpt)

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

    # ------------------------------------------------------------------ coarse -> fine (interface.py:327-380)
    @torch.inference_mode()
    def coarse_to_fine(self, z: torch.Tensor, mask: torch.Tensor = None, return_mask: bool = False, **kwargs):
        assert self.c2f is not None, "No coarse2fine model loaded"
        length = z.shape[-1]
        chunk_len = self.s2t(self.c2f.chunk_size_s)
        n_chunks = math.ceil(z.shape[-1] / chunk_len)
        if length % chunk_len != 0:  # zero pad to a whole number of chunks; padding frames are masked
            pad_len = chunk_len - (length % chunk_len)
            z = torch.nn.functional.pad(z, (0, pad_len))
            mask = torch.nn.functional.pad(mask, (0, pad_len), value=1) if mask is not None else None
        n_append = self.c2f.n_codebooks - z.shape[1]
        if n_append > 0:
            z = torch.cat([z, torch.zeros(z.shape[0], n_append, z.shape[-1], dtype=torch.long, device=z.device)], dim=1)
        if mask is not None:  # conditioning codebooks are never masked
            mask = mask.clone()
            mask[:, :self.c2f.n_conditioning_codebooks, :] = 0
        fine_z = []
        for i in range(n_chunks):
            chunk = z[:, :, i * chunk_len:(i + 1) * chunk_len]
            mask_chunk = mask[:, :, i * chunk_len:(i + 1) * chunk_len] if mask is not None else None
            fine_z.append(self.c2f.generate(codec=self.codec, time_steps=chunk_len, start_tokens=chunk,
                                            return_signal=False, mask=mask_chunk, cfg_guidance=None, **kwargs))
        fine_z = torch.cat(fine_z, dim=-1)
        if return_mask:
            return fine_z[:, :, :length].clone(), pmask.apply_mask(fine_z, mask, self.c2f.mask_token)[0][:, :, :length].clone()
        return fine_z[:, :, :length].clone()

    # ------------------------------------------------------------------ coarse (interface.py:382-452)
    @torch.inference_mode()
    def coarse_vamp(self, z, mask, return_mask=False, gen_fn=None, **kwargs):
        nc = self.coarse.n_codebooks
        cz = z[:, :nc, :].clone()
        mask = mask[:, :nc, :]
        chunk_len = self.s2t(self.coarse.chunk_size_s)
        n_chunks = math.ceil(cz.shape[-1] / chunk_len)
        cz_masked_chunks, cz_vamped_chunks = [], []
        gen_fn = gen_fn or self.coarse.generate
        for i in range(n_chunks):
            chunk = cz[:, :, i * chunk_len:(i + 1) * chunk_len]
            mask_chunk = mask[:, :, i * chunk_len:(i + 1) * chunk_len]
            # first and last frame of a chunk are kept as anchors when the chunk has any unmasked frame, so that
            # stitched chunks do not jump (interface.py:407-413)
            if torch.any(mask_chunk == 0):
                mask_chunk = mask_chunk.clone()
                mask_chunk[:, :, 0] = 0
                mask_chunk[:, :, -1] = 0
            cz_masked_chunk, mask_chunk = pmask.apply_mask(chunk, mask_chunk, self.coarse.mask_token)
            cz_masked_chunks.append(cz_masked_chunk[:, :nc, :])
            cz_vamped_chunks.append(gen_fn(codec=self.codec, time_steps=chunk_len, start_tokens=cz_masked_chunk[:, :nc, :],
                                           return_signal=False, mask=mask_chunk, **kwargs))
        cz_masked = torch.cat(cz_masked_chunks, dim=-1)
        c_vamp = torch.cat(cz_vamped_chunks, dim=-1)
        c_vamp = torch.cat([c_vamp, z[:, nc:, :]], dim=1)  # fine codes ride along untouched
        if return_mask:
            return c_vamp, cz_masked
        return c_vamp

    # ------------------------------------------------------------------ masks (interface.py:454-489)
    def build_mask(self, z: torch.Tensor, sig: AudioSignal = None, rand_mask_intensity: float = 1.0,
                   prefix_s: float = 0.0, suffix_s: float = 0.0, periodic_prompt: int = 7,
                   periodic_prompt_width: int = 1, onset_mask_width: int = 0, _dropout: float = 0.0,
                   upper_codebook_mask: int = 3, ncc: int = 0):
        mask = pmask.linear_random(z, rand_mask_intensity)
        mask = pmask.mask_and(mask, pmask.inpaint(z, self.s2t(prefix_s), self.s2t(suffix_s)))
        mask = pmask.mask_and(mask, pmask.periodic_mask(z, periodic_prompt, periodic_prompt_width, random_roll=True))
        if onset_mask_width > 0:
            assert sig is not None, "must provide a signal to use onset mask"
            mask = pmask.mask_and(mask, pmask.onset_mask(sig, z, self, width=onset_mask_width))
        mask = pmask.dropout(mask, _dropout)
        mask = pmask.codebook_unmask(mask, ncc)
        mask = pmask.codebook_mask(mask, int(upper_codebook_mask), None)
        return mask

    # ------------------------------------------------------------------ vamp (interface.py:491-562)
    def vamp(self, codes: torch.Tensor, mask: torch.Tensor, batch_size: int = 1, feedback_steps: int = 1,
             time_stretch_factor: int = 1, return_mask: bool = False, **kwargs):
        z = codes.expand(batch_size, -1, -1)
        mask = mask.expand(batch_size, -1, -1)
        if time_stretch_factor > 1:  # new in-between frames are always masked (interface.py:510-516)
            z = z.repeat_interleave(time_stretch_factor, dim=-1)
            mask = mask.repeat_interleave(time_stretch_factor, dim=-1)
            added = torch.ones_like(mask)
            added[:, :, ::time_stretch_factor] = 0
            mask = (mask.bool() | added.bool()).long()
        zv = z
        for i in range(feedback_steps):
            zv, mask_z = self.coarse_vamp(zv, mask=mask, return_mask=True, **kwargs)
            mask_z = mask_z.roll(shifts=(i + 1) % feedback_steps, dims=-1)
        if zv.shape[1] < z.shape[1]:
            zv = torch.cat([zv, z[:, self.coarse.n_codebooks:, :]], dim=1)
        # the c2f stage always runs 2 sampling steps with default temperature (interface.py:545-551)
        zv, fine_zv_mask = self.coarse_to_fine(zv, mask=mask, typical_filtering=True, _sampling_steps=2,
                                               return_mask=True)
        mask_z = torch.cat([mask_z[:, :self.coarse.n_codebooks, :], fine_zv_mask[:, self.coarse.n_codebooks:, :]], dim=1)
        if return_mask:
            return zv, mask_z.cpu()
        return zv
