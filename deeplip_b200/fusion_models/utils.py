"""Scoring side of models/fusion_models/utils.py (and its near-duplicate
models/audio_models/utils.py): eer_cos_* (:251-283), *_scorefusion (:331-435),
*_featurefusion (:437-522), feature_normalize (:524-527).

The reference re-reads two .npy files and calls sklearn once PER TRIAL; here the embedding table is
resident on the GPU and one kernel (dl_cosine_score_trials) scores the whole list through the
(enrol_idx, test_idx) gather.  EER stays on the CPU with the very same sklearn / scipy calls
(:280-282), so EER parity reduces to score parity.
"""
import os
import numpy as np
import torch

from .. import ops
from ..trials import TrialList


def eer_from_scores(y_true, y_pred):
    """models/fusion_models/utils.py:280-282."""
    from sklearn.metrics import roc_curve
    from scipy.optimize import brentq
    from scipy.interpolate import interp1d
    fpr, tpr, threshold = roc_curve(y_true, y_pred, pos_label=1)
    eer = brentq(lambda x: 1. - x - interp1d(fpr, tpr)(x), 0., 1.)
    threshold = interp1d(fpr, threshold)(eer)
    return eer, threshold


def feature_normalize(data):
    """utils.py:524-527 on a (N,D) table: biased std per row, no eps -- on the device."""
    z = ops.znorm_concat(data, data, biased=True)
    return z[:, :data.shape[1]].contiguous()


def _dev(x, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(device)


def score_trials(emb, trials, device='cuda', lo=None, hi=None):
    """scores[i] = cos(emb[enrol[i]], emb[test[i]]) for trial lines [lo, hi) -> torch f32 (on device)."""
    sl = slice(lo, hi)
    return ops.cosine_score_trials(_dev(emb, device).float(), _dev(trials.enrol_idx[sl], device),
                                   _dev(trials.test_idx[sl], device))


def score_trials_fusion(emb_audio, emb_video, trials, device='cuda'):
    """0.5 cos(audio) + 0.5 cos(video) (utils.py:343-344, 372-377)."""
    return ops.score_fusion_trials(_dev(emb_audio, device).float(), _dev(emb_video, device).float(),
                                   _dev(trials.enrol_idx, device), _dev(trials.test_idx, device))


def score_trials_featurefusion(emb_audio, emb_video, trials, device='cuda'):
    """biased z-norm per modality, hstack((video, audio)), cosine (utils.py:465-472)."""
    fused = ops.znorm_concat(_dev(emb_audio, device).float(), _dev(emb_video, device).float(), biased=True,
                             video_first=True)
    return ops.cosine_score_trials(fused, _dev(trials.enrol_idx, device), _dev(trials.test_idx, device))


class DenseScorer:
    """The dense formulation north_star names: L2-normalise every embedding once, one tensor-core GEMM
    S = enrol x test^T over the UNIQUE left / right utterances of the list (dl_conv_igemm_bf16, bf16 operands,
    fp32 accumulate), then gather scores[i] = S[row(i), col(i)] (dl_gather_scores).  It does
    n_left * n_right * D MACs for n_trials useful dot products (about 15 000x more arithmetic than the list needs
    on GRID), so `score_trials` (gather-dot, HBM-bound) is the default; this variant exists for lists that are
    dense in enrol x test.  The list-dependent index work (unique utterances, row / column of every trial) is done
    once here on the host; `score(emb)` is device work only."""

    def __init__(self, trials, device='cuda'):
        left, rows = np.unique(trials.enrol_idx, return_inverse=True)
        right, cols = np.unique(trials.test_idx, return_inverse=True)
        self.n_left, self.n_right = len(left), len(right)
        self.n_right_pad = (self.n_right + 7) // 8 * 8
        self.left = _dev(left.astype(np.int64), device)
        right_pad = np.concatenate([right, np.full(self.n_right_pad - self.n_right, right[-1])]).astype(np.int64)
        self.right = _dev(right_pad, device)          # pad columns repeat the last row; never gathered
        self.rows = _dev(rows.astype(np.int32), device)
        self.cols = _dev(cols.astype(np.int32), device)
        self.device = device

    def flop(self, D):
        return 2.0 * self.n_left * self.n_right * D

    def score(self, emb, return_parts=False):
        e = _dev(emb, self.device).float()
        D = e.shape[1]
        assert D % 64 == 0, 'dense scoring needs the embedding dim to be a multiple of 64'
        _, eb = ops.l2_normalize(e, want_bf16=True)
        A = eb.index_select(0, self.left)             # (n_left, D) bf16
        Bm = eb.index_select(0, self.right)           # (n_right_pad, D) bf16  == "weights"
        _, S = ops.conv_igemm(A.view(self.n_left, 1, 1, D), Bm, D, self.n_right_pad, want_bf16=False, want_f32=True)
        out = ops.gather_scores(S, self.rows, self.cols)
        return (out, A, Bm) if return_parts else out


def score_trials_dense(emb, trials, device='cuda'):
    """One-shot form of DenseScorer."""
    return DenseScorer(trials, device).score(emb)


def eer_cos(trials, emb, device='cuda'):
    s = score_trials(emb, trials, device)
    return eer_from_scores(trials.labels, s.cpu().numpy())


# ---------------------------------------------------------------- on-disk hand-off (reference wire format)
def utt_to_relpath(utt):
    return utt.replace('.wav', '.npy')


def save_embeddings(root, utts, emb):
    """One (1,D) float32 .npy per utterance at <root>/<dirname(utt)>/<basename>.npy
    (train_fusion.py:414-417)."""
    emb = emb.detach().cpu().numpy() if torch.is_tensor(emb) else np.asarray(emb)
    for u, e in zip(utts, emb):
        path = os.path.join(root, utt_to_relpath(u))
        os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
        np.save(path, e.reshape(1, -1).astype(np.float32))


def load_embeddings(root, utts):
    """Read each utterance's .npy ONCE (the reference reads two per trial, utils.py:276-277)."""
    return np.stack([np.load(os.path.join(root, utt_to_relpath(u))).reshape(-1) for u in utts]).astype(np.float32)


def save_embedding_table(path, utts, emb):
    """Packed hand-off: ONE (N_utt, D) float32 .npy + a sidecar utterance list, instead of
    one file per utterance.  `path` without extension; writes <path>.npy and <path>.utts.txt."""
    emb = emb.detach().cpu().numpy() if torch.is_tensor(emb) else np.asarray(emb)
    assert emb.ndim == 2 and emb.shape[0] == len(utts)
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    np.save(path + '.npy', np.ascontiguousarray(emb, dtype=np.float32))
    with open(path + '.utts.txt', 'w') as f:
        f.write('\n'.join(utts) + '\n')


def load_embedding_table(path, utts=None):
    """Inverse of save_embedding_table; with `utts` given, rows are returned in that order (a KeyError names a
    missing utterance)."""
    emb = np.load(path + '.npy')
    with open(path + '.utts.txt') as f:
        have = [l.rstrip('\n') for l in f if l.strip()]
    assert emb.shape[0] == len(have), 'table / utterance list length mismatch'
    if utts is None:
        return have, emb
    index = {u: i for i, u in enumerate(have)}
    return list(utts), emb[[index[u] for u in utts]]


def eer_cos_table(trial_path, table_path, device='cuda'):
    """eer_cos_* on the packed table: parse trials, reorder the table to the trial list's utterance order,
    score on the GPU, EER on the CPU."""
    trials = TrialList.from_file(trial_path)
    _, emb = load_embedding_table(table_path, trials.utts)
    return eer_cos(trials, emb, device)


def _eer_cos_dir(exp_dir, sub, trial_path, root='exp', device='cuda'):
    trials = TrialList.from_file(trial_path)
    emb = load_embeddings(os.path.join(root, str(exp_dir), sub), trials.utts)
    return eer_cos(trials, emb, device)


def eer_cos_grid(exp_dir, trial_path='data/data_audio/trial_grid_2w.txt', root='exp', device='cuda'):
    """Reference signature eer_cos_grid(exp_dir) -> (eer, threshold) (utils.py:268-283)."""
    return _eer_cos_dir(exp_dir, 'test_em_grid', trial_path, root, device)


def eer_cos_lomgrid(exp_dir, trial_path='data/data_audio/trial_lomgrid_2w.txt', root='exp', device='cuda'):
    """utils.py:251-266."""
    return _eer_cos_dir(exp_dir, 'test_em_lomgrid', trial_path, root, device)


def eer_plda(trials, emb, classifier, device='cuda'):
    """Per-trial body of eer_plda_grid / eer_plda_lomgrid (models/audio_models/utils.py:285-329), batched: PLDA
    same/different log-likelihood ratios on the GPU (deeplip_b200.plda.Classifier), EER on the CPU."""
    e = _dev(emb, device)
    en = torch.from_numpy(trials.enrol_idx).to(e.device)
    te = torch.from_numpy(trials.test_idx).to(e.device)
    return eer_from_scores(trials.labels, classifier.score_trials(e, en, te).cpu().numpy())


def eer_plda_grid(exp_dir, classifier, trial_path='data/trial/A_grid_trial_2w', root='exp', device='cuda'):
    """eer_plda_grid(exp_dir) of the reference, with the fitted classifier passed in instead of joblib-loaded from
    exp/plda.pkl (the third-party `plda` pickle cannot be read here); embeddings as the per-utterance .npy files of
    exp/<exp_dir>/test_xv_grid (utils.py:316-317)."""
    trials = TrialList.from_file(trial_path)
    emb = load_embeddings(os.path.join(root, str(exp_dir), 'test_xv_grid'), trials.utts)
    return eer_plda(trials, emb, classifier, device)


def eer_plda_lomgrid(exp_dir, classifier, trial_path='data/trial/A_lomgrid_trial_2w', root='exp', device='cuda'):
    """utils.py:285-305."""
    trials = TrialList.from_file(trial_path)
    emb = load_embeddings(os.path.join(root, str(exp_dir), 'test_xv_lomgrid'), trials.utts)
    return eer_plda(trials, emb, classifier, device)
