"""NumPy restatement of the reference's PLDA scorer.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED: the arithmetic lives in the third-party ``plda`` package (github.com/RaviSoji/plda; imported as
``plda`` in train_audio.py, unpinned, absent from /root/reference and from this image).  Call sites:
``train_audio.py:298-341`` (``plda.Classifier().fit_model(embeddings, labels, n_principal_components=20)``,
``joblib.dump``) and ``models/audio_models/utils.py:285-329`` (``eer_plda_grid`` / ``eer_plda_lomgrid``: per trial
``model.transform(em, from_space='D', to_space='U_model')`` then
``model.calc_same_diff_log_likelihood_ratio(U_datum_0, U_datum_1)``).  What follows restates that package's
published algorithm (Ioffe, "Probabilistic Linear Discriminant Analysis", ECCV 2006, as implemented there):

  D --PCA(n_principal_components)--> X --(x - m) A^-1--> U --relevant dims (Psi > 0)--> U_model

  fit: S_b, S_w (class-size weighted, biased covariances); W = generalised eigenvectors of (S_b, S_w);
       Lambda_b = W' S_b W, Lambda_w = W' S_w W; n = N / K;
       A = W^-T (n / (n - 1) diag Lambda_w)^(1/2);  Psi = max(0, (n - 1) / n * diag Lambda_b / diag Lambda_w - 1 / n)
  log marginal likelihood of a set of n vectors of one class, per dimension (prior N(0, psi), class noise N(0, 1)):
       -n/2 log 2pi - 1/2 log(n psi + 1) - 1/2 sum u^2 + 1/2 n^2 psi mean(u)^2 / (n psi + 1)
  same/different log-likelihood ratio of a pair: logp({a, b}) - logp({a}) - logp({b}).
"""
import numpy as np


def scatter_matrices(X, y):
    labels = np.unique(y)
    m = X.mean(axis=0)
    N = X.shape[0]
    n_k = np.array([(y == k).sum() for k in labels], dtype=np.float64)
    m_k = np.stack([X[y == k].mean(axis=0) for k in labels])
    cov_k = np.stack([np.cov(X[y == k].T, bias=True) for k in labels])
    d = m_k - m
    S_b = (d.T * (n_k / N)) @ d
    S_w = (cov_k * (n_k / N)[:, None, None]).sum(axis=0)
    return S_b, S_w


def fit(embeddings, labels, n_principal_components=20):
    """plda.Classifier.fit_model -> plda.Model.__init__: PCA then optimize_maximum_likelihood.  Returns a dict."""
    from scipy.linalg import eigh
    from sklearn.decomposition import PCA
    D = np.asarray(embeddings, dtype=np.float64)
    y = np.asarray(labels)
    # the package calls PCA(n_components=k): sklearn's 'auto' solver turns RANDOMISED for inputs of this size
    # (random_state=None: the reference's own fit is not reproducible run to run); the exact factorisation is pinned
    # here as its deterministic limit
    pca = PCA(n_components=n_principal_components, svd_solver='full')
    pca.fit(D)
    X = pca.transform(D)
    m = X.mean(axis=0)
    S_b, S_w = scatter_matrices(X, y)
    _, W = eigh(S_b, S_w)
    Lb = W.T @ S_b @ W
    Lw = W.T @ S_w @ W
    n = X.shape[0] / float(len(np.unique(y)))
    A = np.linalg.inv(W.T) * np.sqrt(n / (n - 1.0) * np.diag(Lw))
    inv_A = np.linalg.inv(A)
    psi = (n - 1.0) / n * np.diag(Lb) / np.diag(Lw) - 1.0 / n
    psi[psi <= 0] = 0.0
    relevant = np.nonzero(psi > 0)[0]
    return dict(pca_mean=pca.mean_, pca_components=pca.components_, m=m, A=A, inv_A=inv_A, psi=psi,
                relevant=relevant)


def transform_D_to_U_model(model, D):
    X = (np.asarray(D, dtype=np.float64) - model['pca_mean']) @ model['pca_components'].T
    U = (X - model['m']) @ model['inv_A'].T
    return U[..., model['relevant']]


def logp_marginal(model, U_model):
    """plda.Model.calc_logp_marginal_likelihood: U_model (n, R) vectors assumed to share one class."""
    psi = model['psi'][model['relevant']]
    n = U_model.shape[-2]
    npsi1 = n * psi + 1.0
    logc = -0.5 * n * np.log(2.0 * np.pi) - 0.5 * np.log(npsi1)
    e1 = -0.5 * np.sum(U_model ** 2, axis=-2)
    mean = U_model.mean(axis=-2)
    e2 = 0.5 * (n ** 2 * psi * mean ** 2) / npsi1
    return np.sum(logc + e1 + e2, axis=-1)


def same_diff_llr(model, u_p, u_g):
    """plda.Model.calc_same_diff_log_likelihood_ratio on two (1, R) data."""
    same = logp_marginal(model, np.concatenate([u_p, u_g]))
    return same - (logp_marginal(model, u_p) + logp_marginal(model, u_g))


def plda_scores_loop(model, emb, enrol_idx, test_idx):
    """models/audio_models/utils.py:289-305: the per-trial loop of eer_plda_*."""
    out = []
    for a, b in zip(enrol_idx, test_idx):
        em = np.array([emb[a], emb[b]])
        U = transform_D_to_U_model(model, em)
        out.append(same_diff_llr(model, U[0][None, ], U[1][None, ]))
    return np.asarray(out)
