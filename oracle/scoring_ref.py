"""Restatement of the reference's trial scoring + EER.  TEST INFRASTRUCTURE ONLY.

The reference module (models/fusion_models/utils.py) cannot be imported (it needs
``kaldiio``), so its loops are restated here around the very same sklearn / scipy
calls.  PINNED by tests/golden/scoring_*.npz (generated with these calls on seeded
embeddings) and by the sha256 of the two shipped trial lists.
"""
import numpy as np


def parse_trials(path):
    """models/fusion_models/utils.py:271-275: 'label utt1 utt2' per line, rstrip
    removes the trailing TAB / space.  Returns (labels int, [(utt1, utt2)])."""
    labels, pairs = [], []
    with open(path, 'r') as f:
        for line in f:
            line = line.rstrip()
            if not line:
                continue
            lab, u1, u2 = line.split(' ')
            labels.append(int(eval(lab)))
            pairs.append((u1, u2))
    return np.asarray(labels, dtype=np.int64), pairs


def utterance_table(pairs):
    """Deterministic replacement for ``list(set(utts))``
    (models/fusion_models/datasets.py:284-289): first-appearance order
    scanning utt1 then utt2 on every line."""
    table, index = [], {}
    for u1, u2 in pairs:
        for u in (u1, u2):
            if u not in index:
                index[u] = len(table)
                table.append(u)
    enrol = np.asarray([index[a] for a, _ in pairs], dtype=np.int32)
    test = np.asarray([index[b] for _, b in pairs], dtype=np.int32)
    return table, enrol, test


def cosine_scores_loop(emb, enrol_idx, test_idx):
    """models/fusion_models/utils.py:276-279: one sklearn cosine_similarity per
    trial on (1,D) rows.  Returns list of (1,) arrays exactly like y_pred."""
    from sklearn.metrics.pairwise import cosine_similarity
    out = []
    for a, b in zip(enrol_idx, test_idx):
        out.append(cosine_similarity(emb[a].reshape(1, -1), emb[b].reshape(1, -1)).reshape(-1))
    return out


def cosine_scores_vec(emb, enrol_idx, test_idx):
    """Vectorised float64 equivalent (sklearn normalises rows; zero norm -> /1)."""
    e = emb.astype(np.float64)
    n = np.sqrt((e * e).sum(1))
    n[n == 0] = 1.0
    e = e / n[:, None]
    return (e[enrol_idx] * e[test_idx]).sum(1)


def torch_cosine_eps(a, b, eps=1e-8):
    """F.cosine_similarity(a, b, dim=0, eps=1e-8) as used for the video side of
    score fusion (models/fusion_models/utils.py:372)."""
    na = max(float(np.sqrt((a * a).sum())), eps)
    nb = max(float(np.sqrt((b * b).sum())), eps)
    return float((a * b).sum() / (na * nb))


def feature_normalize_np(x):
    """models/fusion_models/utils.py:524-527: biased np.std, no eps."""
    return (x - np.mean(x, axis=0)) / np.std(x, axis=0)


def featurefusion_embedding(audio, video):
    """models/fusion_models/utils.py:465-471: biased z-norm per vector,
    hstack((video, audio))."""
    return np.hstack((feature_normalize_np(video.reshape(-1)), feature_normalize_np(audio.reshape(-1))))


def eer_from_scores(y_true, y_pred):
    """models/fusion_models/utils.py:280-282."""
    from sklearn.metrics import roc_curve
    from scipy.optimize import brentq
    from scipy.interpolate import interp1d
    fpr, tpr, thr = roc_curve(y_true, y_pred, pos_label=1)
    eer = brentq(lambda x: 1. - x - interp1d(fpr, tpr)(x), 0., 1.)
    return eer, interp1d(fpr, thr)(eer)
