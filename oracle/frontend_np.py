"""NumPy restatement of the audio front end the reference runs on DataLoader
workers.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED: the reference calls the third-party package
``python_speech_features`` (unpinned, absent from /root/reference and from this
image).  Call sites: models/fusion_models/datasets.py:227-246 (``_extract_feature``),
:214-215 (``_normalize``); the same body is repeated in
models/audio_models/datasets.py:65-83, 206-224, 305-324.  What follows restates
that package's published v0.6 algorithm for the arguments the reference passes:
``mfcc(data, rate, winlen, winstep, numcep)``, ``fbank(..., nfilt)``,
``logfbank(..., nfilt)``; everything else is the library default
(nfilt=26, nfft=512 at 16 kHz/25 ms, preemph=0.97, ceplifter=22,
appendEnergy=True, rectangular window).  The ``stft`` feature
(datasets.py:237-241) goes through ``librosa.stft`` / ``librosa.magphase``
(unpinned, absent as well): restated from librosa's published algorithm --
periodic Hann window of ``win_length`` centred in ``n_fft``, ``center=True``
padding of ``n_fft // 2`` samples (``reflect`` before librosa 0.10, zeros
since; the reference dates from the ``reflect`` era, which is the default
here), frames every ``hop_length`` samples, rfft, ``log1p(|S|)``.
"""
import math
import numpy as np
from scipy.fftpack import dct

EPS = np.finfo(float).eps


def _round_half_up(x):
    return int(math.floor(x + 0.5))


def calc_nfft(rate, winlen):
    n = 1
    while n < winlen * rate:
        n *= 2
    return n


def preemphasis(sig, coeff=0.97):
    sig = np.asarray(sig, dtype=np.float64)
    return np.concatenate([sig[:1], sig[1:] - coeff * sig[:-1]])


def num_frames(nsamp, frame_len=400, frame_step=160):
    if nsamp <= frame_len:
        return 1
    return 1 + int(math.ceil((nsamp - frame_len) / float(frame_step)))


def frame_signal(sig, frame_len, frame_step):
    frame_len = _round_half_up(frame_len)
    frame_step = _round_half_up(frame_step)
    n = num_frames(len(sig), frame_len, frame_step)
    padlen = (n - 1) * frame_step + frame_len
    pad = np.concatenate([sig, np.zeros(padlen - len(sig))])
    idx = np.arange(frame_len)[None, :] + frame_step * np.arange(n)[:, None]
    return pad[idx]  # rectangular window: multiply by ones


def power_spectrum(frames, nfft):
    return np.square(np.abs(np.fft.rfft(frames, nfft))) / nfft


def hz2mel(hz):
    return 2595.0 * np.log10(1.0 + hz / 700.0)


def mel2hz(mel):
    return 700.0 * (10.0 ** (mel / 2595.0) - 1.0)


def mel_bin_edges(nfilt, nfft, rate, lowfreq=0.0, highfreq=None):
    highfreq = highfreq or rate / 2
    pts = np.linspace(hz2mel(lowfreq), hz2mel(highfreq), nfilt + 2)
    return np.floor((nfft + 1) * mel2hz(pts) / rate)


def mel_filterbank(nfilt, nfft, rate, lowfreq=0.0, highfreq=None):
    b = mel_bin_edges(nfilt, nfft, rate, lowfreq, highfreq)
    fb = np.zeros((nfilt, nfft // 2 + 1))
    for j in range(nfilt):
        for i in range(int(b[j]), int(b[j + 1])):
            fb[j, i] = (i - b[j]) / (b[j + 1] - b[j])
        for i in range(int(b[j + 1]), int(b[j + 2])):
            fb[j, i] = (b[j + 2] - i) / (b[j + 2] - b[j + 1])
    return fb


def fbank(sig, rate=16000, winlen=0.025, winstep=0.01, nfilt=26, nfft=None,
          preemph=0.97):
    nfft = nfft or calc_nfft(rate, winlen)
    sig = preemphasis(sig, preemph)
    frames = frame_signal(sig, winlen * rate, winstep * rate)
    ps = power_spectrum(frames, nfft)
    energy = ps.sum(axis=1)
    energy = np.where(energy == 0, EPS, energy)
    feat = ps @ mel_filterbank(nfilt, nfft, rate).T
    feat = np.where(feat == 0, EPS, feat)
    return feat, energy


def logfbank(sig, rate=16000, winlen=0.025, winstep=0.01, nfilt=26):
    feat, _ = fbank(sig, rate, winlen, winstep, nfilt)
    return np.log(feat)


def lifter(cep, L=22):
    n = np.arange(cep.shape[1])
    return cep * (1.0 + (L / 2.0) * np.sin(np.pi * n / L))


def mfcc(sig, rate=16000, winlen=0.025, winstep=0.01, numcep=13, nfilt=26,
         ceplifter=22, append_energy=True):
    feat, energy = fbank(sig, rate, winlen, winstep, nfilt)
    feat = dct(np.log(feat), type=2, axis=1, norm='ortho')[:, :numcep]
    feat = lifter(feat, ceplifter)
    if append_energy:
        feat[:, 0] = np.log(energy)
    return feat


def stft_logmag(sig, n_fft=512, hop=160, win_length=400, pad_mode='reflect'):
    """datasets.py:237-241: librosa.stft(data, n_fft, hop_length, win_length) -> magphase -> log1p,
    transposed to (T, n_fft // 2 + 1)."""
    sig = np.asarray(sig, dtype=np.float64)
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win_length) / win_length)   # get_window('hann', fftbins=True)
    lpad = (n_fft - win_length) // 2
    win = np.concatenate([np.zeros(lpad), win, np.zeros(n_fft - win_length - lpad)])     # util.pad_center
    y = np.pad(sig, n_fft // 2, mode=pad_mode)
    n = 1 + (len(y) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n)[:, None]
    return np.log1p(np.abs(np.fft.rfft(y[idx] * win[None, :], n_fft, axis=1)))


def delta(feat, N):
    """python_speech_features.delta: sum_n n (f[t+n] - f[t-n]) / (2 sum_n n^2), edges repeated."""
    denom = 2.0 * sum(i * i for i in range(1, N + 1))
    padded = np.pad(feat, ((N, N), (0, 0)), mode='edge')
    w = np.arange(-N, N + 1, dtype=np.float64)
    return np.stack([w @ padded[t:t + 2 * N + 1] for t in range(feat.shape[0])]) / denom


def delta_stack(feat, order=2):
    """models/fusion_models/datasets.py:217-225 (`_delta`): [feat, delta(feat, 1), delta(feat, 2)] side by side."""
    parts = [feat] + [delta(feat, n) for n in range(1, order + 1)]
    return np.hstack(parts)


def cmvn(feat):
    """models/fusion_models/datasets.py:214-215 (biased std, +2e-12)."""
    return (feat - feat.mean(axis=0)) / (feat.std(axis=0) + 2e-12)


def extract_feature(sig, rate=16000, feat_type='mfcc', opts=None):
    """models/fusion_models/datasets.py:227-246; returns (T, F) float32
    (the datasets then transpose to (F, T), :268, :376)."""
    o = dict(win_len=0.025, win_shift=0.01, num_cep=24, num_bin=26, n_fft=512,
             normalize=True, delta=False, pad_mode='reflect')
    o.update(opts or {})
    if feat_type == 'mfcc':
        f = mfcc(sig, rate, o['win_len'], o['win_shift'], numcep=o['num_cep'])
    elif feat_type == 'fbank':
        f, _ = fbank(sig, rate, o['win_len'], o['win_shift'], nfilt=o['num_bin'])
    elif feat_type == 'logfbank':
        f = logfbank(sig, rate, o['win_len'], o['win_shift'], nfilt=o['num_bin'])
    elif feat_type == 'stft':
        f = stft_logmag(sig, o['n_fft'], int(rate * o['win_shift']), int(rate * o['win_len']), o['pad_mode'])
    else:
        raise NotImplementedError("Other features are not implemented!")
    if o['normalize']:
        f = cmvn(f)
    if o['delta']:
        f = delta_stack(f, 2 if o['delta'] is True else int(o['delta']))
    return f.astype(np.float32)
