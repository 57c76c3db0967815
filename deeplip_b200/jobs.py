"""Whole-list jobs: extract every utterance of a trial list, gather the table, score the list, EER.

This is the B200 form of the reference's two-stage test procedure run back to back:
`Trainer.extract_test_xv_grid` / `extract_test_xv_lomgrid` (train_fusion.py:369-420: one utterance at a time,
one `.npy` per utterance) followed by `eer_cos_grid` / `eer_cos_lomgrid` (models/fusion_models/utils.py:251-283:
two `np.load`s + one sklearn call per trial).  SURVEY 8(e) defines the multi-GPU shape of it:

  1. the utterance table of the list (first-appearance order, deeplip_b200.trials) is cut into `world` equal
     contiguous shards; rank r extracts its shard in batches -- no data-path collective;
  2. every batch's fused rows (K8) are written STRAIGHT into the rank's slice of the `(world * per, D)` table,
     which is also the collective's receive buffer, so there is no copy before the gather;
  3. ONE `all_gather_into_tensor` (NCCL, in place) completes the table on every rank
     (GRID 25 834 x 1024 f32 = 105.8 MB, LomGRID 3 541 x 1024 = 14.5 MB);
  4. rank r scores trial lines [r n/world, (r+1) n/world) with the gather-dot kernel, the 20 000 scores are
     gathered, and the EER is computed on the CPU with the reference's own sklearn / scipy calls.

Device-agnostic plumbing (gloo/CPU in the unit tests with an injected `extract`), CUDA kernels in production.
"""
import time

import numpy as np
import torch
import torch.distributed as dist

from . import dist as dl_dist


def _bits_checksum(rows):
    """Order-independent exact checksum of fp32 rows: int64 sum of their bit patterns."""
    return rows.contiguous().view(torch.int32).to(torch.int64).sum()


class _Timer:
    """CUDA events on CUDA tensors' current stream, perf_counter on the CPU (gloo tests)."""

    def __init__(self, device):
        self.cuda = torch.device(device).type == 'cuda'
        self.device = device
        self.marks = []

    def mark(self, name):
        if self.cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(self.device))
            self.marks.append((name, ev))
        else:
            self.marks.append((name, time.perf_counter()))

    def spans_ms(self):
        if self.cuda:
            torch.cuda.synchronize(self.device)
        out = {}
        for (_, a), (n, b) in zip(self.marks[:-1], self.marks[1:]):
            out[n] = a.elapsed_time(b) if self.cuda else (b - a) * 1e3
        return out


class TrialListJob:
    """One pass over one trial list on `world` ranks.

    extract(idx_lo, idx_hi, out_rows): write the fused embeddings of utterances [idx_lo, idx_hi) of
    `trials.utts` into `out_rows` ((idx_hi - idx_lo, D) f32, a slice of the table).  In production this is
    `AVExtractor.extract(..., out=out_rows)` around the caller's loader."""

    def __init__(self, trials, dim, rank=0, world=1, device='cuda', global_batch=256):
        self.trials, self.dim, self.rank, self.world = trials, dim, rank, world
        self.device = torch.device(device)
        self.n_utts = len(trials.utts)
        self.per = dl_dist.shard_size(self.n_utts, world)             # rows per rank, shards padded to equal size
        self.lo, self.hi = dl_dist.shard_range(self.n_utts, rank, world)
        if global_batch % world:
            raise ValueError('global batch %d does not divide over %d ranks' % (global_batch, world))
        self.batch = global_batch // world
        # the table IS the all-gather buffer; pad rows (only the last shard has any) stay zero
        self.table = torch.zeros((self.per * world, dim), dtype=torch.float32, device=self.device)
        self.local = self.table[rank * self.per:(rank + 1) * self.per]
        sl = trials.shard(rank, world)
        self._sync_token = torch.zeros((1,), dtype=torch.float32, device=self.device)
        self.enrol = torch.from_numpy(np.ascontiguousarray(trials.enrol_idx[sl])).to(self.device)
        self.test = torch.from_numpy(np.ascontiguousarray(trials.test_idx[sl])).to(self.device)

    def batches(self):
        for b0 in range(self.lo, self.hi, self.batch):
            yield b0, min(b0 + self.batch, self.hi)

    def gather_table(self):
        """ONE collective: every rank's (per, D) slice into the full table, in place on NCCL."""
        if self.world == 1:
            return
        send = self.local if dist.get_backend() == 'nccl' else self.local.clone()
        dist.all_gather_into_tensor(self.table, send)

    def verify_gather(self):
        """Every shard of the gathered table carries the bits its owner computed: owners publish the exact
        checksum of their slice taken BEFORE the collective, every rank recomputes all of them AFTER it."""
        if self.world == 1:
            return True
        mine = self._pre_checksum.reshape(1)
        allc = torch.empty((self.world,), dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(allc, mine)
        now = torch.stack([_bits_checksum(self.table[r * self.per:(r + 1) * self.per]) for r in range(self.world)])
        ok = torch.tensor([int(torch.equal(now, allc))], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        return bool(ok.item())

    def run(self, extract, score, eer_fn=None):
        """score(table, enrol_idx, test_idx) -> scores f32 (device).  Returns a dict with the scores (full list,
        every rank), the EER (rank 0) and the per-phase times in ms (max over ranks is the caller's business)."""
        tm = _Timer(self.device)
        dl_dist.barrier()
        tm.mark('start')
        for b0, b1 in self.batches():
            extract(b0, b1, self.local[b0 - self.lo:b1 - self.lo])
        tm.mark('extract')
        self._pre_checksum = _bits_checksum(self.local) if self.world > 1 else None
        tm.mark('checksum')
        if self.world > 1:        # a one-element all_reduce on the stream: the ranks' skew (unequal clocks, a slower
            dist.all_reduce(self._sync_token)      # shard) lands in its own span, not in the all_gather's
        tm.mark('rank_skew')
        self.gather_table()
        tm.mark('all_gather')
        local_scores = score(self.table[:self.n_utts], self.enrol, self.test)
        tm.mark('score')
        scores = dl_dist.gather_scores(local_scores, len(self.trials), self.rank, self.world)
        tm.mark('gather_scores')
        spans = tm.spans_ms()
        out = {'ms': spans, 'scores': scores, 'n_utts': self.n_utts, 'n_trials': len(self.trials),
               'per_rank_rows': self.per, 'per_gpu_batch': self.batch,
               'allgather_bytes': self.per * self.world * self.dim * 4 if self.world > 1 else 0}
        if eer_fn is not None and self.rank == 0:
            t0 = time.perf_counter()
            out['eer'], out['threshold'] = eer_fn(self.trials.labels, scores.detach().cpu().numpy())
            out['eer_ms_cpu'] = (time.perf_counter() - t0) * 1e3
        return out
