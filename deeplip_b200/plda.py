"""PLDA trial scoring on the GPU: drop-in for the `plda.Classifier` the reference trains in
train_audio.py:298-341 and scores with in models/audio_models/utils.py:285-329 (eer_plda_grid / eer_plda_lomgrid).

The third-party `plda` package is unpinned and absent, so parity is UNPINNED: this follows its
published algorithm (PCA -> maximum-likelihood PLDA of Ioffe 2006 -> same/different log-likelihood ratio from the
marginal likelihoods).  Fitting is a few dense float64 factorisations on a (N_dev, D) matrix and stays on the host
like the reference's (it is training, not the hot path); the hot path -- transform every utterance, score every
trial -- runs in two kernels of libdeeplip_b200.so with no per-trial host work.

    clf = Classifier(); clf.fit_model(dev_embeddings, dev_labels, n_principal_components=20)
    llr = clf.score_trials(emb_cuda, enrol_idx_cuda, test_idx_cuda)          # (n_trials,) f32
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .ops import _need_cuda, _ptr, _stream


class Classifier:
    """Same entry point as plda.Classifier (train_audio.py:339-340): fit_model(X, Y, n_principal_components)."""

    def __init__(self):
        self.model = None

    def fit_model(self, X, Y, n_principal_components=None):
        from scipy.linalg import eigh
        D = np.asarray(X, dtype=np.float64)
        y = np.asarray(Y)
        assert D.ndim == 2 and y.shape == (D.shape[0],)
        k = min(D.shape) if n_principal_components is None else int(n_principal_components)
        # PCA: exact SVD of the centred data (sklearn's PCA, which the package calls, switches to a randomised
        # solver for large inputs; the exact factorisation is its deterministic limit)
        mean = D.mean(axis=0)
        _, _, Vt = np.linalg.svd(D - mean, full_matrices=False)
        comp = Vt[:k]
        Xp = (D - mean) @ comp.T
        m = Xp.mean(axis=0)
        labels, inv = np.unique(y, return_inverse=True)
        K, N = len(labels), Xp.shape[0]
        counts = np.bincount(inv, minlength=K).astype(np.float64)
        means = np.zeros((K, k))
        np.add.at(means, inv, Xp)
        means /= counts[:, None]
        dev = Xp - means[inv]
        S_w = dev.T @ dev / N                                       # = sum_k n_k cov_k(bias=True) / N
        dm = means - m
        S_b = (dm.T * (counts / N)) @ dm
        _, W = eigh(S_b, S_w)
        lam_b = np.einsum('ij,ik,kj->j', W, S_b, W)
        lam_w = np.einsum('ij,ik,kj->j', W, S_w, W)
        n = N / float(K)
        A = np.linalg.inv(W.T) * np.sqrt(n / (n - 1.0) * lam_w)
        inv_A = np.linalg.inv(A)
        psi = (n - 1.0) / n * lam_b / lam_w - 1.0 / n
        psi[psi <= 0] = 0.0
        rel = np.nonzero(psi > 0)[0]
        if len(rel) == 0 or len(rel) > 32:
            raise RuntimeError('PLDA: %d relevant dimensions (the scoring kernels take 1..32)' % len(rel))
        # u_model = x @ Wt + b  (D -> X -> U -> relevant dims, folded into one affine map)
        Wt = (comp.T @ inv_A.T)[:, rel]
        b = -((mean @ comp.T + m) @ inv_A.T)[rel]
        p = psi[rel]
        self.model = dict(M=np.ascontiguousarray(Wt.T), bias=b, psi=p,
                          k1=p / (2.0 * (2.0 * p + 1.0)), k2=p / (2.0 * (p + 1.0)),
                          c0=float(np.sum(np.log(p + 1.0) - 0.5 * np.log(2.0 * p + 1.0))))
        self._dev = {}
        return self

    def _device_params(self, device):
        key = str(device)
        if key not in self._dev:
            f = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32)).contiguous().to(device)
            mo = self.model
            self._dev[key] = (f(mo['M']), f(mo['bias']), f(mo['k1']), f(mo['k2']))
        return self._dev[key]

    @torch.no_grad()
    def transform(self, emb):
        """emb (N_utt, D) f32 CUDA -> U_model (N_utt, R) f32: `model.transform(em, 'D', 'U_model')` for every row."""
        if self.model is None:
            raise RuntimeError('PLDA: fit_model first')
        _need_cuda(emb)
        emb = emb.contiguous().float()
        M, bias, _, _ = self._device_params(emb.device)
        R, D = M.shape
        if emb.shape[1] != D:
            raise RuntimeError('PLDA: embeddings have %d dimensions, the model was fitted on %d' % (emb.shape[1], D))
        u = torch.empty((emb.shape[0], R), device=emb.device, dtype=torch.float32)
        _lib.check(_lib.lib().dl_plda_transform(_ptr(emb), emb.shape[0], D, _ptr(M), _ptr(bias), R, _ptr(u), _stream()),
                   'dl_plda_transform')
        return u

    @torch.no_grad()
    def score_trials(self, emb, enrol_idx, test_idx):
        """Same/different log-likelihood ratio of every trial (models/audio_models/utils.py:296-304, batched)."""
        _need_cuda(emb, enrol_idx, test_idx)
        assert enrol_idx.dtype == torch.int32 and test_idx.dtype == torch.int32 and enrol_idx.numel() == test_idx.numel()
        u = self.transform(emb)
        _, _, k1, k2 = self._device_params(emb.device)
        n = enrol_idx.numel()
        scores = torch.empty((n,), device=emb.device, dtype=torch.float32)
        _lib.check(_lib.lib().dl_plda_llr_trials(_ptr(u), u.shape[0], u.shape[1], _ptr(k1), _ptr(k2),
                                                 C.c_float(self.model['c0']), _ptr(enrol_idx.contiguous()),
                                                 _ptr(test_idx.contiguous()), n, _ptr(scores), _stream()),
                   'dl_plda_llr_trials')
        return scores
