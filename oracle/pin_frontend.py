#!/usr/bin/env python
"""Pin the front-end oracle against the REAL third-party packages.  TEST INFRASTRUCTURE ONLY.

    pip install python_speech_features==0.6 librosa          # on any machine with an index
    python oracle/pin_frontend.py                            # writes tests/golden/frontend.npz
    python -m pytest tests/test_oracle.py -k frontend_golden # compares oracle/frontend_np.py with it

The reference computes its audio features with `python_speech_features` (`mfcc` / `fbank` / `logfbank` / `delta`)
and `librosa` (`stft` + `magphase`), both unpinned and absent from /root/reference and from the build image
(models/fusion_models/datasets.py:6-7, 229-241, 217-225).  oracle/frontend_np.py restates their published
algorithms, so its parity is UNPINNED until this script has been run where the packages exist.  The script feeds
the packages the same seeded inputs the tests use (deeplip_b200.synth.speech_like_audio, plus the edge lengths the
framing arithmetic cares about), through the reference's own call expressions, and stores inputs' seeds, outputs
(float64, before CMVN) and the package versions.  `tests/test_oracle.py::test_frontend_golden_from_real_packages`
loads the file when present and skips (saying "parity unpinned") when it is not.

This file imports nothing from /root/reference: the call expressions below are the reference's, restated.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

RATE = 16000
# (name, speakers, nsamp, seed): GRID shape, a length that is an exact number of hops, one that is not, a short one
CASES = [('grid3s', [1, 2, 3], 48000, 1), ('exact_hops', [4], 400 + 160 * 50, 2), ('ragged', [5, 6], 33333, 3),
         ('short', [7], 4001, 4)]


def reference_calls(sig, psf, librosa):
    """The five feature variants exactly as models/fusion_models/datasets.py:229-241 and :217-225 call them."""
    out = {}
    out['mfcc'] = psf.mfcc(sig, RATE, winlen=0.025, winstep=0.01, numcep=24)                       # :229
    out['fbank'] = psf.fbank(sig, RATE, winlen=0.025, winstep=0.01, nfilt=24)[0]                   # :231-232
    out['logfbank'] = psf.logfbank(sig, RATE, winlen=0.025, winstep=0.01, nfilt=60)                # :233-234
    S = librosa.stft(sig.astype(np.float32), n_fft=512, hop_length=160, win_length=400)            # :237-239
    mag, _ = librosa.magphase(S)
    out['stft'] = np.log1p(mag).T                                                                  # :240-241
    d1 = psf.delta(out['mfcc'], 1)                                                                 # _delta, :217-225
    d2 = psf.delta(out['mfcc'], 2)
    out['mfcc_delta'] = np.hstack((out['mfcc'], d1, d2))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'tests', 'golden', 'frontend.npz'))
    args = ap.parse_args()
    try:
        import python_speech_features as psf
        import librosa
    except ImportError as e:
        sys.exit('pin_frontend: %s -- install python_speech_features==0.6 and librosa first (this image has neither, '
                 'so the oracle stays "parity unpinned" here)' % e)
    from deeplip_b200 import synth
    blob = {'psf_version': np.array(getattr(psf, '__version__', '0.6?')), 'librosa_version': np.array(librosa.__version__),
            'case_names': np.array([c[0] for c in CASES])}
    for name, spk, nsamp, seed in CASES:
        wav = synth.speech_like_audio(spk, nsamp=nsamp, seed=seed)
        blob[name + '/spk'] = np.array(spk)
        blob[name + '/nsamp'] = np.array(nsamp)
        blob[name + '/seed'] = np.array(seed)
        blob[name + '/wav_crc'] = np.array(int(np.frombuffer(wav.tobytes(), dtype=np.uint32).sum(dtype=np.uint64)))
        for i in range(len(spk)):
            for kind, val in reference_calls(wav[i].astype(np.float64), psf, librosa).items():
                blob['%s/%d/%s' % (name, i, kind)] = np.asarray(val, dtype=np.float64)
    np.savez_compressed(args.out, **blob)
    print('wrote', args.out, 'with', len(blob), 'arrays; psf', blob['psf_version'], 'librosa', blob['librosa_version'])


if __name__ == '__main__':
    main()
