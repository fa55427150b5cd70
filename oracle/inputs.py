"""TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Synthetic MOSI/MOSEI-shaped batches, numpy-seeded so the same
batch is produced on any machine.

BERT layout follows multimodal_driver.py:143-173 (`prepare_bert_input`):
  ids   = [CLS=101, tokens..., SEP=102, PAD=0 ...]; mask = 1 on the first l positions;
  segment ids all 0; visual/acoustic rows are zero at CLS, SEP and pad positions.
XLNet layout follows multimodal_driver.py:176-205 (`prepare_xlnet_input`):
  left-padded with PAD=5; [tokens..., SEP=4, CLS=3] at the end; mask 0 on pads;
  segment ids 3 (pad) / 0 (tokens+SEP) / 2 (CLS); modality rows zero at SEP, CLS, pad.
"""
from __future__ import annotations

import numpy as np
import torch


def bert_batch(batch=48, seq=50, visual_dim=47, acoustic_dim=74, vocab=30522, seed=1234,
               min_len=5, full_length=False, lens=None):
    rng = np.random.Generator(np.random.PCG64(seed))
    ids = np.zeros((batch, seq), np.int64)
    mask = np.zeros((batch, seq), np.int64)
    tt = np.zeros((batch, seq), np.int64)
    vis = rng.standard_normal((batch, seq, visual_dim), dtype=np.float32)
    aco = rng.standard_normal((batch, seq, acoustic_dim), dtype=np.float32)
    if lens is not None:                       # explicit lengths (edge cases: l = 2 is CLS + SEP only)
        lens = np.asarray(lens, np.int64)
        assert lens.shape == (batch,) and lens.min() >= 2 and lens.max() <= seq
    else:
        lens = np.full(batch, seq) if full_length else rng.integers(min_len, seq + 1, size=batch)
    for b in range(batch):
        l = int(lens[b])
        ids[b, 0] = 101
        ids[b, 1:l - 1] = rng.integers(1000, vocab, size=l - 2)
        ids[b, l - 1] = 102
        mask[b, :l] = 1
        vis[b, 0] = 0; vis[b, l - 1:] = 0
        aco[b, 0] = 0; aco[b, l - 1:] = 0
    labels = rng.uniform(-3, 3, size=(batch,)).astype(np.float32)
    return dict(input_ids=torch.from_numpy(ids), visual=torch.from_numpy(vis),
                acoustic=torch.from_numpy(aco), attention_mask=torch.from_numpy(mask),
                token_type_ids=torch.from_numpy(tt), labels=torch.from_numpy(labels))


def xlnet_batch(batch=48, seq=50, visual_dim=47, acoustic_dim=74, vocab=32000, seed=1234,
                min_len=5, full_length=False, lens=None):
    rng = np.random.Generator(np.random.PCG64(seed))
    ids = np.full((batch, seq), 5, np.int64)
    mask = np.zeros((batch, seq), np.int64)
    tt = np.full((batch, seq), 3, np.int64)
    vis = rng.standard_normal((batch, seq, visual_dim), dtype=np.float32)
    aco = rng.standard_normal((batch, seq, acoustic_dim), dtype=np.float32)
    if lens is not None:                       # explicit lengths (edge cases: l = 2 is CLS + SEP only)
        lens = np.asarray(lens, np.int64)
        assert lens.shape == (batch,) and lens.min() >= 2 and lens.max() <= seq
    else:
        lens = np.full(batch, seq) if full_length else rng.integers(min_len, seq + 1, size=batch)
    for b in range(batch):
        l = int(lens[b])  # tokens + SEP + CLS
        pad = seq - l
        ids[b, pad:seq - 2] = rng.integers(10, vocab, size=l - 2)
        ids[b, seq - 2] = 4
        ids[b, seq - 1] = 3
        mask[b, pad:] = 1
        tt[b, pad:seq - 1] = 0
        tt[b, seq - 1] = 2
        vis[b, :pad] = 0; vis[b, seq - 2:] = 0
        aco[b, :pad] = 0; aco[b, seq - 2:] = 0
    labels = rng.uniform(-3, 3, size=(batch,)).astype(np.float32)
    return dict(input_ids=torch.from_numpy(ids), visual=torch.from_numpy(vis),
                acoustic=torch.from_numpy(aco), attention_mask=torch.from_numpy(mask),
                token_type_ids=torch.from_numpy(tt), labels=torch.from_numpy(labels))


def mag_case(batch=4, seq=50, hidden=768, visual_dim=47, acoustic_dim=74, seed=7, zero_rows=False):
    """Config 1 (BASELINE.json configs[0]): standalone MAG inputs."""
    rng = np.random.Generator(np.random.PCG64(seed))
    text = rng.standard_normal((batch, seq, hidden), dtype=np.float32)
    vis = rng.standard_normal((batch, seq, visual_dim), dtype=np.float32)
    aco = rng.standard_normal((batch, seq, acoustic_dim), dtype=np.float32)
    if zero_rows:
        vis[:, ::3] = 0
        aco[:, ::3] = 0
    return torch.from_numpy(text), torch.from_numpy(vis), torch.from_numpy(aco)
