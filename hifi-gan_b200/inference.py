"""Batched drop-in for the reference's inference drivers (src/inference.py:37-62 wav -> mel -> wav,
src/inference_e2e.py:34-57 mel -> wav) — SURVEY.md §8(f) rank 1.

Same command lines (`--input_wavs_dir / --input_mels_dir / --output_dir / --checkpoint_file`, config.json beside
the checkpoint, src/inference.py:68-80), same output file names and the same int16 samples.  What changes is the
schedule: the reference runs batch 1 per file with a blocking device-to-host copy of fp32 audio per file; here files
are sorted by length and stacked into length-bucketed, RAGGED batches: the kernels take per-item lengths and treat
the rows past an item's end as the zero padding it would see alone (hg_conv1d_fwd's item_len), so every file's
samples stay bit-identical to the per-file call while a directory of different-length utterances still fills the
GPU.  The `* MAX_WAV_VALUE -> int16` conversion runs on the device (hg_float_to_int16) and the int16 batch crosses to
pinned host memory in one copy.

    python -m hifigan_b200.inference      --checkpoint_file cp/g_02500000 [--input_wavs_dir test_files]
    python -m hifigan_b200.inference e2e  --checkpoint_file cp/g_02500000 [--input_mels_dir test_mel_files]
"""
from __future__ import annotations

import argparse
import json
import os
from collections import defaultdict
from typing import Dict, Iterable, List, Tuple

import numpy as np
import torch

from . import _lib
from .env import AttrDict
from .meldataset import MAX_WAV_VALUE, mel_spectrogram
from .models import Generator
from .utils import load_checkpoint


def audio_to_int16(y: torch.Tensor) -> torch.Tensor:
    """fp32 waveform (any shape) on the GPU -> int16 with the reference's `audio * 32768 -> astype('int16')` rule."""
    y = y.contiguous().float()
    out = torch.empty(y.shape, dtype=torch.int16, device=y.device)
    _lib.check(_lib.lib().hg_float_to_int16(y.data_ptr(), y.numel(), MAX_WAV_VALUE, out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "hg_float_to_int16")
    return out


def bucket_by_length(items: Iterable[Tuple[str, torch.Tensor]], max_batch: int,
                     max_waste: float = 0.25) -> List[List[Tuple[str, torch.Tensor]]]:
    """Length-bucketed batches of (name, mel [80, F]) items: sorted by frame count, a batch takes up to max_batch
    neighbours as long as the padding it implies stays below max_waste of the batch (the Generator computes the padded
    rows too; `lengths` keeps every item's samples bit-identical to a one-file call).  max_waste = 0: equal lengths only."""
    ordered = sorted(items, key=lambda it: int(it[1].shape[-1]))
    batches: List[List[Tuple[str, torch.Tensor]]] = []
    cur: List[Tuple[str, torch.Tensor]] = []
    total = 0
    for name, mel in ordered:
        f = int(mel.shape[-1])
        if cur and (len(cur) >= max_batch or 1.0 - (total + f) / (f * (len(cur) + 1)) > max_waste):
            batches.append(cur)
            cur, total = [], 0
        cur.append((name, mel))
        total += f
    if cur:
        batches.append(cur)
    return batches


@torch.no_grad()
def vocode(generator: Generator, batches, sampling_rate: int, output_dir: str, suffix: str, writer=None) -> List[str]:
    """Run the bucketed batches and write `<name><suffix>.wav` int16 files; returns the paths in processing order."""
    if writer is None:
        from scipy.io.wavfile import write as writer
    written = []
    dev = next(generator.parameters()).device
    for batch in batches:
        frames = [int(m.shape[-1]) for _, m in batch]
        fmax = max(frames)
        if min(frames) == fmax:
            x = torch.stack([m for _, m in batch]).to(dev)
            y = generator(x)
        else:                                               # ragged batch: zero-padded mels + per-item lengths
            x = torch.zeros(len(batch), batch[0][1].shape[0], fmax, dtype=torch.float32, device=dev)
            for i, (_, m) in enumerate(batch):
                x[i, :, : m.shape[-1]] = m.to(dev)
            y = generator(x, lengths=torch.tensor(frames, dtype=torch.int32, device=dev))
        audio = audio_to_int16(y.squeeze(1))                                 # [B, T] int16 on the device
        host = torch.empty(audio.shape, dtype=torch.int16).pin_memory()
        host.copy_(audio, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        hop = audio.shape[1] // fmax
        for (name, _), row, f in zip(batch, host.numpy(), frames):
            row = row[: f * hop]
            path = os.path.join(output_dir, os.path.splitext(name)[0] + suffix + ".wav")
            writer(path, sampling_rate, row)
            print(path)
            written.append(path)
    return written


def _load_generator(checkpoint_file: str, h, device) -> Generator:
    generator = Generator(h).to(device)
    state_dict_g = load_checkpoint(checkpoint_file, device)
    generator.load_state_dict(state_dict_g['generator'])
    generator.eval()
    generator.remove_weight_norm()
    return generator


def inference(a, h, device) -> List[str]:
    """src/inference.py:37-62: every wav of a.input_wavs_dir -> mel (fmax of the config) -> Generator -> wav."""
    from .meldataset import load_wav
    generator = _load_generator(a.checkpoint_file, h, device)
    os.makedirs(a.output_dir, exist_ok=True)
    items = []
    for name in os.listdir(a.input_wavs_dir):
        wav, sr = load_wav(os.path.join(a.input_wavs_dir, name))
        wav = torch.as_tensor(np.asarray(wav), dtype=torch.float32) / MAX_WAV_VALUE   # reference quirk kept (:51-52)
        mel = mel_spectrogram(wav.reshape(1, -1).to(device), h.n_fft, h.num_mels, h.sampling_rate, h.hop_size,
                              h.win_size, h.fmin, h.fmax)
        items.append((name, mel[0]))
    return vocode(generator, bucket_by_length(items, a.max_batch), h.sampling_rate, a.output_dir, "_generated")


def inference_e2e(a, h, device) -> List[str]:
    """src/inference_e2e.py:34-57: every .npy mel of a.input_mels_dir -> Generator -> wav."""
    generator = _load_generator(a.checkpoint_file, h, device)
    os.makedirs(a.output_dir, exist_ok=True)
    items = []
    for name in os.listdir(a.input_mels_dir):
        x = torch.as_tensor(np.load(os.path.join(a.input_mels_dir, name)), dtype=torch.float32)
        items.append((name, x.reshape(-1, x.shape[-2], x.shape[-1])[0]))          # [1,80,F] or [80,F] files
    return vocode(generator, bucket_by_length(items, a.max_batch), h.sampling_rate, a.output_dir, "_generated_e2e")


def main(argv=None) -> None:
    print('Initializing Inference Process..')
    parser = argparse.ArgumentParser()
    parser.add_argument('mode', nargs='?', default='wav', choices=['wav', 'e2e'])
    parser.add_argument('--input_wavs_dir', default='test_files')
    parser.add_argument('--input_mels_dir', default='test_mel_files')
    parser.add_argument('--output_dir', default=None)
    parser.add_argument('--checkpoint_file', required=True)
    parser.add_argument('--max_batch', type=int, default=64, help='files stacked per Generator call (length-bucketed)')
    a = parser.parse_args(argv)
    if a.output_dir is None:
        a.output_dir = 'generated_files' if a.mode == 'wav' else 'generated_files_from_mel'
    config_file = os.path.join(os.path.split(a.checkpoint_file)[0], 'config.json')
    with open(config_file) as f:
        h = AttrDict(json.loads(f.read()))
    torch.manual_seed(h.seed)
    if not torch.cuda.is_available():
        raise RuntimeError("hifigan_b200 has no CPU path: a B200 is required")
    torch.cuda.manual_seed(h.seed)
    device = torch.device('cuda')
    (inference if a.mode == 'wav' else inference_e2e)(a, h, device)


if __name__ == '__main__':
    main()
