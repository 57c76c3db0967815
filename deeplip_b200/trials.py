"""Trial-list parsing and the deterministic utterance table.

Reference: models/fusion_models/utils.py:271-275 (one 'label utt1 utt2' line per trial, rstrip
drops the trailing TAB/space) and models/fusion_models/datasets.py:284-289, where the utterance
set is `list(set(...))` -- hash-randomised per process.  Here the utterance index is
first-appearance order scanning utt1 then utt2 on every line, so (enrol_idx, test_idx) are stable.
Integer work on the host; bit-exact by construction and checked against the oracle.
"""
import numpy as np


class TrialList:
    def __init__(self, labels, pairs):
        self.labels = np.asarray(labels, dtype=np.int64)
        self.pairs = pairs
        table, index = [], {}
        enrol = np.empty(len(pairs), dtype=np.int32)
        test = np.empty(len(pairs), dtype=np.int32)
        for i, (u1, u2) in enumerate(pairs):
            for u in (u1, u2):
                if u not in index:
                    index[u] = len(table)
                    table.append(u)
            enrol[i] = index[u1]
            test[i] = index[u2]
        self.utts = table
        self.index = index
        self.enrol_idx = enrol
        self.test_idx = test

    def __len__(self):
        return len(self.pairs)

    @classmethod
    def from_file(cls, path):
        labels, pairs = [], []
        with open(path, 'r') as f:
            for line in f:
                line = line.rstrip()
                if not line:
                    continue
                parts = line.split(' ')
                if len(parts) != 3 or parts[0] not in ('0', '1'):
                    raise ValueError('bad trial line %r (want "<0|1> <utt1> <utt2>")' % line)
                labels.append(int(parts[0]))
                pairs.append((parts[1], parts[2]))
        return cls(labels, pairs)

    def shard(self, rank, world):
        """Contiguous slice of trial lines scored by `rank`."""
        n = len(self)
        per = -(-n // world)
        return slice(min(rank * per, n), min((rank + 1) * per, n))
