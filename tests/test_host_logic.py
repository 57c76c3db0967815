"""CPU suite: host-side logic (trial parsing, packing, sharding), the C-ABI library loads and exports
every symbol include/deeplip_b200.h declares, no compute call succeeds without a GPU, and the
multi-rank plumbing works on gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_build_and_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from deeplip_b200 import _lib
    hdr = re.sub(r'/\*.*?\*/', '', open(os.path.join(ROOT, 'include', 'deeplip_b200.h')).read(), flags=re.S)
    declared = set(re.findall(r'\b(dl_\w+)\s*\(', hdr))
    assert len(declared) >= 17
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert set(_lib.SIGNATURES) <= declared
    assert _lib.lib().dl_version() >= 100


def test_argument_validation_without_touching_the_gpu():
    from deeplip_b200 import _lib
    l = _lib.lib()
    assert l.dl_stat_pool(None, 1, 1, 8, 8, None, None, None, 0, None) == -1
    assert b'null' in l.dl_last_error()
    d = _lib.ConvDesc(1, 4, 4, 12, 12, 8, 1, 1, 1, 1, 0, 0, 1, 1, 8, 8, 1.0)
    one = ctypes.c_void_p(16)
    assert l.dl_conv_igemm_bf16(one, one, one, one, one, None, one, None, None, None, ctypes.byref(d), None) == -1
    assert b'ldx' in l.dl_last_error()
    d2 = _lib.ConvDesc(1, 8, 8, 64, 64, 512, 3, 3, 2, 2, 1, 1, 1, 1, 256, 512, 1.0)
    d2.split_channel, d2.y_split = 128, 16                     # sibling-conv split must fall on a 256-channel boundary
    assert l.dl_conv_igemm_bf16(one, one, one, one, one, None, one, None, None, None, ctypes.byref(d2), None) == -1
    assert b'split_channel' in l.dl_last_error()
    d2.split_channel, d2.y_split = 256, 0                      # ... and needs the second output
    assert l.dl_conv_igemm_bf16(one, one, one, one, one, None, one, None, None, None, ctypes.byref(d2), None) == -1
    assert b'y_split' in l.dl_last_error()
    assert l.dl_frontend_features(one, None, 1, 48000, 0, 24, 1, 0, None, 64, one, 298, None) == -1
    assert b'299' in l.dl_last_error()
    assert l.dl_frontend_features(one, None, 1, 48000, 3, 257, 1, 0, None, 320, one, 299, None) == -1   # stft framing
    assert b'301' in l.dl_last_error()
    assert l.dl_frontend_features(one, None, 1, 48000, 3, 24, 1, 0, None, 64, one, 301, None) == -1
    assert b'257' in l.dl_last_error()
    assert l.dl_frontend_features(one, None, 1, 48000, 0, 24, 1, 3, None, 128, one, 299, None) == -1      # delta order
    assert b'delta' in l.dl_last_error()


def test_fft512_phase_functions_on_cpu(tmp_path):
    """The warp FFT of the front end (deeplip_b200/csrc/fft512.cuh) is written as host+device phase functions:
    run them on the CPU, lane by lane between the kernel's __syncwarp points, against a direct DFT."""
    exe = str(tmp_path / 'fft512_host_check')
    subprocess.run(['g++', '-O2', '-o', exe, os.path.join(ROOT, 'tests', 'fft512_host_check.cpp')], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    assert out[0] == 'max_err' and float(out[1]) < 2e-4 and float(out[2]) < 1e-11, out


@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU-only behaviour')
def test_no_cpu_fallback():
    from deeplip_b200 import ops
    from deeplip_b200.pipeline import build_models
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.znorm_concat(torch.randn(2, 8), torch.randn(2, 8))
    audio, video = build_models('cpu')
    with pytest.raises(RuntimeError):
        video(torch.zeros(1, 1, 2, 88, 88), lengths=[2])
    with pytest.raises(RuntimeError):
        audio.extract_embedding(torch.zeros(1, 24, 100))


@pytest.mark.parametrize('feat_type,F,nsamp,lengths,pad,delta', [
    ('mfcc', 24, 16000, None, 'reflect', 0), ('logfbank', 60, 8000, None, 'reflect', 0),
    ('fbank', 24, 8000, None, 'reflect', 0), ('mfcc', 24, 20000, [20000, 12345, 300], 'reflect', 0),
    ('stft', 257, 8000, None, 'reflect', 0), ('stft', 257, 8000, None, 'constant', 0),
    ('stft', 257, 20000, [20000, 12345, 700], 'reflect', 0),
    ('mfcc', 24, 20000, [20000, 12345, 300], 'reflect', 2), ('logfbank', 60, 8000, None, 'reflect', 2),
    ('mfcc', 13, 6000, None, 'reflect', 1), ('stft', 257, 6000, [6000, 1000], 'reflect', 2)])
def test_frontend_kernel_source_on_cpu_threads(tmp_path, feat_type, F, nsamp, lengths, pad, delta):
    """The generation-2 front-end kernels are compiled FROM THEIR CUDA SOURCE for the CPU (one OS thread per CUDA
    thread, tests/frontend_cpu_emul.cpp) and compared with the oracle: framing, FFT, mel/DCT tables, ragged lengths,
    stft padding, CMVN and the bf16 channels-last copy are all checked without a GPU."""
    from oracle import frontend_np
    from deeplip_b200 import synth
    exe = str(tmp_path / 'emul')
    subprocess.run(['g++', '-std=c++20', '-O2', '-pthread', '-Wno-unknown-pragmas', '-Wno-attributes', '-o', exe,
                    os.path.join(ROOT, 'tests', 'frontend_cpu_emul.cpp')], check=True)
    B = 2 if lengths is None else len(lengths)
    wav = synth.speech_like_audio(list(range(B)), nsamp=nsamp, seed=1).astype(np.float32)
    wav.tofile(str(tmp_path / 'wav.f32'))
    lf = '-'
    if lengths is not None:
        lf = str(tmp_path / 'len.i32')
        np.asarray(lengths, np.int32).tofile(lf)
    kind = {'mfcc': 0, 'fbank': 1, 'logfbank': 2, 'stft': 3}[feat_type]
    out = subprocess.run([exe, str(kind), str(F), str(nsamp), str(B), '0' if pad == 'reflect' else '1',
                          str(1 + 10 * delta), str(tmp_path / 'wav.f32'), lf, str(tmp_path / 'out.bin')], check=True,
                         capture_output=True, text=True).stdout.split()
    T, ld = int(out[1]), int(out[3])
    raw = np.fromfile(str(tmp_path / 'out.bin'), dtype=np.uint8)
    Fb, F = F, F * (1 + delta)                           # F: rows of one utterance incl. the delta features
    f32 = raw[:B * F * T * 4].view(np.float32).reshape(B, F, T)
    bf = (raw[B * F * T * 4:].view(np.uint16).reshape(B, T, ld).astype(np.uint32) << 16).view(np.float32)
    for i in range(B):
        n = nsamp if lengths is None else lengths[i]
        ref = frontend_np.extract_feature(wav[i, :n].astype(np.float64), 16000, feat_type,
                                          dict(num_cep=Fb, num_bin=Fb, pad_mode=pad, delta=delta)).T
        assert np.abs(f32[i][:, :ref.shape[1]] - ref).max() < 2e-3
        assert np.all(f32[i][:, ref.shape[1]:] == 0)                     # padding frames of a ragged batch
        assert np.abs(bf[i, :ref.shape[1], :F].T - ref).max() < 5e-2
        assert np.all(bf[i, ref.shape[1]:, :] == 0) and np.all(bf[i, :, F:] == 0)


@pytest.mark.parametrize('is_u8,H,W,Hraw,Wraw,frames', [(1, 88, 88, 96, 96, 2), (1, 88, 88, 90, 90, 2),
                                                        (1, 64, 64, 100, 100, 2), (1, 88, 88, 88, 88, 2),
                                                        (0, 88, 88, 88, 88, 2), (0, 32, 36, 32, 36, 2),
                                                        (1, 88, 88, 96, 96, 4), (0, 32, 32, 32, 32, 4)])
def test_stem_prepass_kernel_source_on_cpu_threads(tmp_path, is_u8, H, W, Hraw, Wraw, frames):
    """stem_prepass2_kernel (V1 fused: /255, centre crop, mean/std, zero border, bf16) run from its CUDA source on
    CPU threads: aligned-word and byte load paths, crop offsets that make the word base negative, f32 input."""
    exe = str(tmp_path / 'emul')
    subprocess.run(['g++', '-std=c++20', '-O2', '-pthread', '-Wno-unknown-pragmas', '-Wno-attributes', '-o', exe,
                    os.path.join(ROOT, 'tests', 'frontend_cpu_emul.cpp')], check=True)
    rng = np.random.default_rng(0)
    if is_u8:
        x = rng.integers(0, 256, (frames, Hraw, Wraw), dtype=np.uint8)
        dh, dw = (Hraw - H) // 2, (Wraw - W) // 2          # CenterCrop, models/video_models/preprocess.py:88-90
        ref = (x[:, dh:dh + H, dw:dw + W].astype(np.float32) / 255.0 - 0.421) / 0.165
    else:
        x = rng.standard_normal((frames, H, W)).astype(np.float32)
        ref = x
    x.tofile(str(tmp_path / 'in.bin'))
    subprocess.run([exe, 'prepass', str(is_u8), str(frames), str(H), str(W), str(Hraw), str(Wraw),
                    str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')], check=True, capture_output=True)
    rows, pitch = H + 8, (W + 8 + 7) // 8 * 8
    got = (np.fromfile(str(tmp_path / 'out.bin'), np.uint16).reshape(frames, rows, pitch).astype(np.uint32) << 16
           ).view(np.float32)
    exp = np.zeros_like(got)
    exp[:, 3:3 + H, 3:3 + W] = ref
    if frames == 4:          # the harness runs 4 frames as two clips of T = 2 with lengths (2, 1): a ragged batch whose
        exp[3] = 0           # padding frame must enter the stem as NORMALISED zeros (pad_packed_collate convention)
    assert (np.abs(got - exp) / np.maximum(1, np.abs(exp))).max() < 5e-3          # bf16 rounding
    border = np.ones_like(got, dtype=bool)
    border[:, 3:3 + H, 3:3 + W] = False
    assert np.all(got[border] == 0)


@pytest.mark.parametrize('M,C,Cout,wb,ws2', [(64, 3000, 512, 1, 1), (64, 512, 512, 0, 1), (3, 1024, 24, 1, 0),
                                             (100, 512, 8, 1, 1), (33, 72, 24, 1, 1), (150, 64, 8, 1, 1)])
def test_linear_small_kernel_source_on_cpu_threads(tmp_path, M, C, Cout, wb, ws2):
    """linear_small_kernel (fc heads, tdnn.py:89-101) from its CUDA source on CPU threads vs a float64 product of the
    same bf16 operands: K split over 16 warps, 1 / 2 rows per lane, row groups on grid.y, both outputs and epilogues."""
    exe = str(tmp_path / 'emul')
    subprocess.run(['g++', '-std=c++20', '-O2', '-pthread', '-Wno-unknown-pragmas', '-Wno-attributes', '-o', exe,
                    os.path.join(ROOT, 'tests', 'frontend_cpu_emul.cpp')], check=True)
    rng = np.random.default_rng(0)
    ldx, ldw = (C + 63) // 64 * 64 + 8, (C + 63) // 64 * 64
    x = torch.from_numpy(rng.standard_normal((M, ldx)).astype(np.float32)).to(torch.bfloat16)
    w = torch.from_numpy((rng.standard_normal((Cout, ldw)) / np.sqrt(C)).astype(np.float32)).to(torch.bfloat16)
    prm = rng.standard_normal((5, Cout)).astype(np.float32)          # scale, shift, slope, scale2, shift2
    prm[2] = np.abs(prm[2]) * 0.3
    with open(tmp_path / 'in.bin', 'wb') as f:
        f.write(x.view(torch.int16).numpy().tobytes())
        f.write(w.view(torch.int16).numpy().tobytes())
        f.write(prm.tobytes())
    subprocess.run([exe, 'linear', str(M), str(C), str(Cout), str(ldx), str(ldw), str(wb), str(ws2),
                    str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')], check=True)
    raw = np.fromfile(str(tmp_path / 'out.bin'), np.uint8)
    y = (raw[:M * Cout * 2].view(np.uint16).astype(np.uint32) << 16).view(np.float32).reshape(M, Cout)
    yf = raw[M * Cout * 2:].view(np.float32).reshape(M, Cout)
    acc = x[:, :C].double().numpy() @ w[:, :C].double().numpy().T         # columns >= C of x are ignored
    o2 = acc * prm[3] + prm[4] if ws2 else acc
    o2 = np.where(o2 > 0, o2, o2 * 0.2)
    assert np.abs(yf - o2).max() < 1e-4
    if wb:
        v = acc * prm[0] + prm[1]
        v = np.where(v > 0, v, v * prm[2])
        assert (np.abs(y - v) / np.maximum(1, np.abs(v))).max() < 5e-3


def _plda_fixture(dim=128, n_spk=30, per=20, n_trials=400):
    from deeplip_b200 import synth
    rng = np.random.default_rng(0)
    spk = np.repeat(np.arange(n_spk), per)
    X = synth.structured_embeddings(spk.tolist(), dim=dim, seed=3).astype(np.float64)
    test_spk = np.repeat(np.arange(100, 112), 6)
    E = synth.structured_embeddings(test_spk.tolist(), dim=dim, seed=5).astype(np.float32)
    en = rng.integers(0, len(E), n_trials).astype(np.int32)
    te = rng.integers(0, len(E), n_trials).astype(np.int32)
    return X, spk, E, test_spk, en, te


def test_plda_host_fit_and_kernel_source_against_oracle(tmp_path):
    """PLDA: the product's fit (deeplip_b200/plda.py: exact-SVD PCA, folded affine map, closed-form
    LLR constants) against the oracle's restatement of the `plda` package (general marginal likelihoods, per-trial
    loop), and the two scoring kernels run from their CUDA source on CPU threads."""
    from oracle import plda_ref, scoring_ref
    from deeplip_b200.plda import Classifier
    X, spk, E, test_spk, en, te = _plda_fixture()
    mo = plda_ref.fit(X, spk, 20)
    ref = plda_ref.plda_scores_loop(mo, E, en, te)
    # oracle sanity: symmetric in its two arguments, targets score higher, EER is low on speaker-structured data
    assert np.abs(ref - plda_ref.plda_scores_loop(mo, E, te, en)).max() < 1e-9
    lab = (test_spk[en] == test_spk[te]).astype(int)
    assert ref[lab == 1].mean() > ref[lab == 0].mean()
    assert scoring_ref.eer_from_scores(lab, ref)[0] < 0.2
    m = Classifier().fit_model(X, spk, 20).model
    R, D = m['M'].shape
    assert R == len(mo['relevant'])
    u = E.astype(np.float64) @ m['M'].T + m['bias']
    # U is defined up to a sign per dimension (eigenvector signs); the scores do not depend on it
    assert np.abs(np.abs(u) - np.abs(plda_ref.transform_D_to_U_model(mo, E))).max() < 1e-8
    a, b = u[en], u[te]
    closed = m['c0'] + ((a + b) ** 2 * m['k1'] - (a * a + b * b) * m['k2']).sum(1)
    assert np.abs(closed - ref).max() < 1e-9
    # the kernels, from source
    exe = str(tmp_path / 'emul')
    subprocess.run(['g++', '-std=c++20', '-O2', '-pthread', '-Wno-unknown-pragmas', '-Wno-attributes', '-o', exe,
                    os.path.join(ROOT, 'tests', 'frontend_cpu_emul.cpp')], check=True)
    en2, te2 = en.copy(), te.copy()
    en2[7] = len(E)                                     # out-of-range index -> NaN score
    with open(tmp_path / 'in.bin', 'wb') as f:
        for arr in (E, m['M'], m['bias'], m['k1'], m['k2']):
            f.write(np.ascontiguousarray(arr, dtype=np.float32).tobytes())
        f.write(en2.tobytes())
        f.write(te2.tobytes())
    subprocess.run([exe, 'plda', str(len(E)), str(D), str(R), str(len(en)), repr(m['c0']), str(tmp_path / 'in.bin'),
                    str(tmp_path / 'out.bin'), '-', '-'], check=True)
    raw = np.fromfile(str(tmp_path / 'out.bin'), np.float32)
    ug, sg = raw[:len(E) * R].reshape(len(E), R), raw[len(E) * R:]
    assert np.abs(ug - u).max() < 1e-3 * max(1.0, np.abs(u).max())
    assert np.isnan(sg[7])
    ok = np.arange(len(en)) != 7
    assert np.abs(sg[ok] - ref[ok]).max() < 2e-3 * max(1.0, np.abs(ref).max())


def test_trial_list_parsing_matches_oracle(tmp_path):
    from deeplip_b200.trials import TrialList
    from oracle import scoring_ref
    import gpu_checks as G
    for kind in ('grid', 'lomgrid'):
        p = G.make_trial_file(str(tmp_path / ('t_%s.txt' % kind)), kind, n_target=200, n_non=700)
        tl = TrialList.from_file(p)
        labels, pairs = scoring_ref.parse_trials(p)
        table, enrol, test = scoring_ref.utterance_table(pairs)
        assert table == tl.utts
        assert np.array_equal(enrol, tl.enrol_idx) and np.array_equal(test, tl.test_idx)
        assert np.array_equal(labels, tl.labels) and tl.enrol_idx.dtype == np.int32
    # edge cases: empty file, blank lines, malformed line
    e = tmp_path / 'empty.txt'
    e.write_text('\n\n')
    assert len(TrialList.from_file(str(e))) == 0
    b = tmp_path / 'bad.txt'
    b.write_text('2 a b\n')
    with pytest.raises(ValueError):
        TrialList.from_file(str(b))
    ref = '/root/reference/database/trial_grid_v1.txt'
    if os.path.exists(ref):
        tl = TrialList.from_file(ref)
        assert len(tl) == 20000 and len(tl.utts) == 25834 and int(tl.labels.sum()) == 4000
        sh = [tl.shard(r, 8) for r in range(8)]
        assert sh[0] == slice(0, 2500) and sh[7] == slice(17500, 20000)


def test_packed_embedding_table_roundtrip(tmp_path):
    from deeplip_b200.fusion_models import utils as U
    utts = ['s1/a.wav', 's2/b.wav', 's3/c.wav']
    emb = np.arange(12, dtype=np.float32).reshape(3, 4)
    U.save_embedding_table(str(tmp_path / 'tab'), utts, torch.from_numpy(emb))
    have, e = U.load_embedding_table(str(tmp_path / 'tab'))
    assert have == utts and np.array_equal(e, emb)
    _, e2 = U.load_embedding_table(str(tmp_path / 'tab'), ['s3/c.wav', 's1/a.wav'])
    assert np.array_equal(e2, emb[[2, 0]])
    with pytest.raises(KeyError):
        U.load_embedding_table(str(tmp_path / 'tab'), ['nope.wav'])


def test_packing_layouts():
    from deeplip_b200 import packing
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = packing.pack_conv_weight(w)
    assert p.shape == (8, 9 * 64) and p.dtype == torch.bfloat16
    assert float(p[1, (2 * 3 + 1) * 64 + 2]) == float(w[1, 2, 2, 1])           # K = (r*S+s)*64 + c
    assert float(p[2:].abs().max()) == 0 and float(p[0, 3]) == 0
    ws = torch.randn(64, 1, 5, 7, 7)
    ps = packing.pack_stem_weight(ws)
    assert ps.shape == (64, 320)
    assert float(ps[5, 2 * 64 + 3 * 8 + 6]) == float(ws[5, 0, 2, 3, 6].to(torch.bfloat16))
    assert float(ps[:, 7::8].abs().max()) == 0                                   # kw = 7 padding column
    s, h = packing.fold_bn(torch.tensor([2.0]), torch.tensor([0.5]), torch.tensor([1.0]), torch.tensor([3.0]),
                           conv_bias=torch.tensor([0.25]))
    y = (0.7 + 0.25 - 1.0) / (3.0 + 1e-5) ** 0.5 * 2.0 + 0.5
    assert abs(float(0.7 * s + h) - y) < 1e-6


def test_state_dict_keys_match_the_reference_layout():
    from deeplip_b200 import synth
    from deeplip_b200.pipeline import build_models
    audio, video = build_models('cpu')
    vk = set(video.state_dict().keys())
    assert set(synth.make_video_state_dict().keys()) == vk
    assert 'frontend3D.0.weight' in vk and 'trunk.layer2.0.downsample.0.weight' in vk
    assert tuple(video.state_dict()['frontend3D.0.weight'].shape) == (64, 1, 5, 7, 7)
    ak = set(audio.state_dict().keys())
    assert {'tdnn.0.context_layer.weight', 'tdnn.9.bn.running_var', 'fc1.weight', 'bn2.bias'} <= ak
    assert tuple(audio.state_dict()['fc1.weight'].shape) == (512, 3000)
    # DataParallel 'module.' prefix and TCN-head keys of a real checkpoint are tolerated
    sd = {'module.' + k: v for k, v in synth.make_video_state_dict().items()}
    sd['module.tcn.tcn_output.weight'] = torch.zeros(500, 768)
    video.load_state_dict(sd)


_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'tests'))
from deeplip_b200 import dist as D, synth
from deeplip_b200.trials import TrialList
from oracle import scoring_ref
import gpu_checks as G
rank, world, _ = D.init_from_env('gloo')
tl = TrialList.from_file(%(trial)r)
n = len(tl.utts)
full = torch.from_numpy(synth.structured_embeddings([synth.speaker_of_utt(u) for u in tl.utts], dim=32, seed=4))
lo, hi = D.shard_range(n, rank, world)
table = D.all_gather_rows(full[lo:hi].clone(), n, rank, world)            # each rank "extracted" only its shard
assert torch.equal(table, full), 'all_gather_rows mismatch'
sl = tl.shard(rank, world)
loc = torch.from_numpy(scoring_ref.cosine_scores_vec(table.numpy(), tl.enrol_idx[sl], tl.test_idx[sl])).float()
allsc = D.gather_scores(loc, len(tl), rank, world)
ref = scoring_ref.cosine_scores_vec(full.numpy(), tl.enrol_idx, tl.test_idx).astype(np.float32)
assert np.array_equal(allsc.numpy(), ref), 'sharded scores differ from the single-rank scores'
assert D.max_over_ranks(rank + 1.5, 'cpu') == world + 0.5
D.barrier(); dist.destroy_process_group()
sys.stdout.write('rank %%d ok\n' %% rank); sys.stdout.flush()
'''


def test_two_rank_gloo_shard_gather_score(tmp_path):
    import gpu_checks as G
    trial = G.make_trial_file(str(tmp_path / 't.txt'), 'lomgrid', n_target=101, n_non=300)
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER % {'root': ROOT, 'trial': trial})
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr',
           '127.0.0.1', '--master-port', '29611', str(script)]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=240)
    text = out.stdout.decode()
    assert out.returncode == 0, text[-2000:]
    assert 'rank 0 ok' in text and 'rank 1 ok' in text


def test_host_pipeline_rejects_single_slot():
    """one staging slot would be overwritten by the next upload before the kernels read it."""
    from deeplip_b200.pipeline import HostPipeline
    with pytest.raises(ValueError):
        HostPipeline(object(), 'cuda', slots=1)


def test_buffer_cache_keeps_old_shapes_and_pins_recorded():
    """persistent buffers are cached per shape (a new shape must not free what a CUDA graph captured) and
    `recording()` hands the graph owner references that survive eviction."""
    from deeplip_b200.ops import BufferCache
    c = BufferCache(cap=2)
    a = c.get(('k', 1), lambda: torch.zeros(4))
    assert c.get(('k', 1), lambda: torch.ones(4)) is a
    with c.recording() as rec:
        b = c.get(('k', 2), lambda: torch.zeros(8))
        assert c.get(('k', 1), lambda: None) is a
    assert rec[0] is b and rec[1] is a
    c.get(('k', 3), lambda: torch.zeros(1))
    c.get(('k', 4), lambda: torch.zeros(1))          # evicts shapes 2 and 1 from the cache ...
    assert c.get(('k', 2), lambda: torch.ones(8)) is not b
    assert rec[0] is b and float(b.sum()) == 0.0      # ... but the recorded references still own the old buffers


def test_tdnn_rejects_utterance_shorter_than_receptive_field():
    """the reference's Conv1d raises on too-short input; the drop-in must not return NaN embeddings."""
    from deeplip_b200 import synth
    from deeplip_b200.audio_models.tdnn import SpeakerEmbNet
    net = SpeakerEmbNet(synth.audio_opts('etdnn', 'statistic')).eval()
    assert net.min_frames == net.context_loss + 2 == 24
    x = torch.zeros((1, net.context_loss + 1, 64), dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        net.embed_ntc(x)


REAL_LISTS = {'grid': ('trial_grid_v1.txt', '58d2936203e004a82fdef9ee5c6076f6c3812655803d2ad1d84583111577ccab',
                       25834, 20000, 14900),
              'lomgrid': ('trial_lomgrid_v1.txt', '968e12cabba8067ae8662211fe8cea5bad901a66b83625737d32b7181ba61632',
                          3541, 3526, 3447)}


@pytest.mark.parametrize('name', ['grid', 'lomgrid'])
def test_real_trial_lists_index_bit_exact(name):
    """The reference's two shipped trial lists (database/trial_{grid,lomgrid}_v1.txt, reproduced under tests/golden/):
    file sha256, and the WHOLE (enrol_idx, test_idx) vectors of the product parser against the position-weighted
    xor checksums gen_golden.py stored from the oracle's parse of the reference files (bit-exact trial indexing)."""
    import hashlib
    from deeplip_b200.trials import TrialList
    fname, sha, n_utts, n_left, n_right = REAL_LISTS[name]
    path = os.path.join(ROOT, 'tests', 'golden', fname)
    assert hashlib.sha256(open(path, 'rb').read()).hexdigest() == sha
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'scoring.npz'))
    assert str(g[name + '_sha256']) == sha
    tl = TrialList.from_file(path)
    assert len(tl) == 20000 and len(tl.utts) == n_utts == int(g[name + '_n_utts'])
    assert int(tl.labels.sum()) == 4000 == int(g[name + '_labels_sum']) and tl.labels[:4000].all() and not tl.labels[4000:].any()
    assert len(np.unique(tl.enrol_idx)) == n_left and len(np.unique(tl.test_idx)) == n_right
    w = np.arange(len(tl)) + 1
    assert int(np.bitwise_xor.reduce(tl.enrol_idx.astype(np.int64) * w)) == int(g[name + '_enrol_crc'])
    assert int(np.bitwise_xor.reduce(tl.test_idx.astype(np.int64) * w)) == int(g[name + '_test_crc'])
    assert np.array_equal(tl.enrol_idx[:64], g[name + '_enrol_head']) and np.array_equal(tl.test_idx[:64], g[name + '_test_head'])
    # label == (speaker1 == speaker2) on every line
    from deeplip_b200 import synth
    spk = np.array([synth.speaker_of_utt(u) for u in tl.utts])
    assert np.array_equal((spk[tl.enrol_idx] == spk[tl.test_idx]).astype(np.int64), tl.labels)


_JOB_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from deeplip_b200 import dist as D, synth
from deeplip_b200.jobs import TrialListJob
from deeplip_b200.trials import TrialList
from deeplip_b200.fusion_models.utils import eer_from_scores
from oracle import scoring_ref
rank, world, _ = D.init_from_env('gloo')
tl = TrialList.from_file(%(trial)r)
full = torch.from_numpy(synth.structured_embeddings([synth.speaker_of_utt(u) for u in tl.utts], dim=32, seed=4, within=3.0))
job = TrialListJob(tl, 32, rank, world, device='cpu', global_batch=256)
assert job.batch == 128 and job.per == 1771 and job.table.shape == (3542, 32)
calls = []
def extract(lo, hi, out):                 # this rank "extracts" only rows it owns, straight into its table slice
    assert job.lo <= lo < hi <= job.hi and hi - lo <= job.batch and out.shape == (hi - lo, 32)
    assert out.data_ptr() == job.table[rank * job.per + lo - job.lo].data_ptr()
    calls.append((lo, hi)); out.copy_(full[lo:hi])
def score(table, en, te):
    return torch.from_numpy(scoring_ref.cosine_scores_vec(table.numpy(), en.numpy(), te.numpy())).float()
res = job.run(extract, score, eer_fn=eer_from_scores)
assert calls[0][0] == job.lo and calls[-1][1] == job.hi and len(calls) == -(-(job.hi - job.lo) // 128)
assert job.verify_gather()
assert torch.equal(job.table[:len(tl.utts)], full) and float(job.table[len(tl.utts):].abs().sum()) == 0
ref = scoring_ref.cosine_scores_vec(full.numpy(), tl.enrol_idx, tl.test_idx).astype(np.float32)
assert np.array_equal(res['scores'].numpy(), ref)
if rank == 0:
    eer, _ = scoring_ref.eer_from_scores(tl.labels, list(ref.reshape(-1, 1)))
    assert abs(res['eer'] - eer) < 1e-12 and 0.0 < eer < 0.5
    assert set(res['ms']) == {'extract', 'checksum', 'rank_skew', 'all_gather', 'score', 'gather_scores'}
job.table[0, 0] += 1.0                    # a corrupted gather must be caught on every rank
assert not job.verify_gather()
D.barrier(); dist.destroy_process_group()
sys.stdout.write('rank %%d ok\n' %% rank); sys.stdout.flush()
'''


def test_two_rank_gloo_trial_list_job_on_real_lomgrid(tmp_path):
    """configs[3] as a job (deeplip_b200.jobs.TrialListJob) on the real trial_lomgrid_v1 list, world_size 2 on gloo:
    shards, in-table extraction, ONE all-gather, sharded scoring, gathered scores == the single-process oracle, EER."""
    script = tmp_path / 'job_worker.py'
    script.write_text(_JOB_WORKER % {'root': ROOT, 'trial': os.path.join(ROOT, 'tests', 'golden', 'trial_lomgrid_v1.txt')})
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr',
           '127.0.0.1', '--master-port', '29613', str(script)]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=240)
    text = out.stdout.decode()
    assert out.returncode == 0, text[-3000:]
    assert 'rank 0 ok' in text and 'rank 1 ok' in text


def test_stem2_layout_arithmetic_against_torch_on_cpu():
    """The second-generation stem's data layout restated in numpy (tools/stem2_emulate.py): pre-pass frames -> 14-row
    strips -> two parity planes of unfolded rows -> no-swizzle K-major descriptor views (SBO 128 B, LBO = plane; the
    shared window-row-6 step with LBO = one stage) x the weight stack built from packing.pack_stem_weight ->
    accumulator [128 lanes x 176 columns] -> per-thread BN / PReLU / 3x3-s2 max-pool with the carried conv row, compared
    with torch's conv3d + max_pool3d.  Checks the design the CUDA kernel implements; the kernel itself is checked on the
    GPU (tests/gpu_checks.py: stem_case)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'tools', 'stem2_emulate.py')], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'OK' in r.stdout, r.stdout + r.stderr


def test_overlapped_gather_is_the_identity_for_one_rank():
    """dist.OverlappedGather at world == 1 neither needs CUDA nor touches its argument (bench.py's N = 1 line); the
    multi-rank form is checked bit for bit against the synchronous collective on NCCL inside bench.py
    (`step_collective.overlapped_equals_synchronous_all_ranks`)."""
    from deeplip_b200 import dist as D
    og = D.OverlappedGather(7, 0, 1, 'cpu')
    x = torch.arange(14.0).view(7, 2)
    assert og(x) is x and og.stream is None
    og.wait()


@pytest.mark.parametrize('B,T,HW,C,lengths', [(3, 75, 9, 512, [75, 40, 1]), (2, 37, 9, 64, [37, 33]), (2, 5, 4, 72, [5, 2])])
def test_k4_kernels_source_on_cpu_threads(tmp_path, B, T, HW, C, lengths):
    """frame_pool_kernel (spatial mean + masked temporal mean, model.py:16-17 / train_fusion.py:400) and
    temporal_mean_kernel (the temporal half alone, used when the conv epilogue took the spatial mean) from their CUDA
    source on CPU threads: both against numpy, and the two utterance means against each other BIT FOR BIT -- the
    claim dl_conv_desc.avgpool rests on."""
    exe = str(tmp_path / 'emul')
    subprocess.run(['g++', '-std=c++20', '-O2', '-pthread', '-Wno-unknown-pragmas', '-Wno-attributes', '-o', exe,
                    os.path.join(ROOT, 'tests', 'frontend_cpu_emul.cpp')], check=True)
    rng = np.random.default_rng(1)
    xf = rng.standard_normal((B * T, HW, C)).astype(np.float32)
    xb = ((xf.view(np.uint32) + 0x7fff + ((xf.view(np.uint32) >> 16) & 1)) >> 16).astype(np.uint16)      # bf16 rn
    xr = (xb.astype(np.uint32) << 16).view(np.float32)
    with open(str(tmp_path / 'in.bin'), 'wb') as f:
        f.write(np.asarray(lengths, np.int32).tobytes())
        f.write(xb.tobytes())
    subprocess.run([exe, 'pool', str(B), str(T), str(HW), str(C), str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')],
                   check=True, capture_output=True)
    out = np.fromfile(str(tmp_path / 'out.bin'), np.float32)
    ff = out[:B * T * C].reshape(B, T, C)
    um = out[B * T * C:B * T * C + B * C].reshape(B, C)
    um2 = out[B * T * C + B * C:].reshape(B, C)
    ref_ff = xr.reshape(B, T, HW, C).astype(np.float64).mean(2)
    assert np.abs(ff - ref_ff).max() < 1e-5
    for b, n in enumerate(lengths):
        assert np.abs(um[b] - ref_ff[b, :n].mean(0)).max() < 1e-5
    assert np.array_equal(um.view(np.uint32), um2.view(np.uint32))


def test_pool_first_identity_of_the_stem_epilogue():
    """stem2_conv3d.cuh pools the RAW accumulators where a channel's BN + PReLU is monotone: for slope >= 0,
    max_i f(v_i) == f(max_i v_i) when scale >= 0 and == f(min_i v_i) when scale <= 0, f(v) = prelu(fma(v, scale, shift)),
    bit for bit in f32 (every rounding step of f is monotone, and f is applied to one of the v_i either way).  Checked
    here on random 3x3 windows incl. ties, zeros, a zero scale and slopes 0 and > 1; a negative slope breaks it (the
    kernel then takes the general order), which the last assertion demonstrates."""
    rng = np.random.default_rng(0)
    v = rng.standard_normal((4000, 9)).astype(np.float32) * 3
    v[::7, 3] = v[::7, 5]                       # ties
    v[::11] = 0.0                               # all-zero windows (padding frames of a ragged batch)

    def f(x, sc, sh, sl):
        z = (x.astype(np.float64) * np.float64(sc) + np.float64(sh)).astype(np.float32)     # fma: one rounding
        return np.where(z > 0, z, (z * np.float32(sl)).astype(np.float32))
    for sc, sh, sl in [(1.3, -0.2, 0.25), (-0.7, 0.4, 0.0), (0.0, 0.3, 0.5), (2.5, 0.0, 1.7), (-1.1, -0.3, 0.9)]:
        general = f(v, sc, sh, sl).max(axis=1)
        pooled = f(v.max(axis=1) if sc >= 0 else v.min(axis=1), sc, sh, sl)
        assert np.array_equal(general.view(np.uint32), pooled.view(np.uint32)), (sc, sh, sl)
    general = f(v, 1.3, -0.2, -0.5).max(axis=1)
    assert not np.array_equal(general, f(v.max(axis=1), 1.3, -0.2, -0.5))
