"""-m gpu: kernel-level parity, CUDA path (through the C ABI) vs the CPU oracle."""
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available(), reason='needs a CUDA device (no CPU fallback)')]

if torch.cuda.is_available():
    import gpu_checks as G
else:
    G = None


@pytest.mark.parametrize('name', ['l1_3x3_64', 'l2_3x3s2_64_128', 'l2_1x1s2_64_128', 'l2_3x3_128', 'l3_3x3s2_128_256',
                                  'l3_3x3_256', 'l4_3x3s2_256_512', 'l4_3x3_512', 'tdnn_k5_24_512', 'tdnn_k3d3_512',
                                  'tdnn_k1_512_1500', 'fc_3000_512', 'fc_3000_512_igemm', 'fc_3000_512_b64',
                                  'fc_512_512_b100', 'fc_512_512_b300', 'fc_1024_512_b33_ld', 'conv1x1_small_map', 'many_tiles', 'pair_128', 'pair_256_s2', 'pair_512_odd',
                                  'pair_tdnn_k1_1504', 'pair_tdnn_k3_res_odd_rows'])
def test_conv_igemm(name):
    G.conv_case(**G.CONV_CASES[name])


def test_conv_center_only_hint_is_bit_identical_on_the_pair_kernel():
    G.conv_center_only_case()


def test_frontend_int16_pcm_input_equals_float_input_bit_for_bit():
    G.frontend_pcm16_case()


def test_staged_epilogue_equals_per_lane_stores_bit_for_bit():
    G.staged_epilogue_bitwise_case()


@pytest.mark.parametrize('kw', [dict(), dict(N=3, H=8, W=40, residual=False), dict(N=1, H=3, W=8), dict(N=300)])
def test_conv3x3_halo(kw):
    G.halo_case(**kw)


@pytest.mark.parametrize('kw', [
    dict(N=3, H=11, W=11, C=128, Cout=128, R=3, S=3, pad=(1, 1), guard=(1, 1), residual=True),      # 1-CTA kernel
    dict(N=150, H=11, W=11, C=128, Cout=128, R=3, S=3, pad=(1, 1), guard=(1, 1), residual=True),    # CTA pairs
    dict(N=600, H=6, W=6, C=256, Cout=256, R=3, S=3, pad=(1, 1), guard=(1, 1)),
    dict(N=2, H=5, W=9, C=64, Cout=64, R=3, S=3, pad=(2, 2), dil=(2, 2), guard=(2, 3), residual=True),
    dict(N=2, H=1, W=140, C=512, Cout=512, R=1, S=3, dil=(1, 3)),                                   # TDNN, no guards
    dict(N=80, H=1, W=300, C=512, Cout=512, R=1, S=5, dil=(1, 1)),
    dict(N=64, H=1, W=296, C=512, Cout=1504, R=1, S=1),
])
def test_conv_igemm_guarded_linear(kw):
    G.conv_lin_case(**kw)


@pytest.mark.parametrize('kw', [dict(), dict(N=2, H=11, W=11, C=128, Cout=256, R=1, S=1, pad=0)])
def test_conv_igemm_guarded_io(kw):
    G.conv_guarded_io_case(**kw)


def test_conv_igemm_rejects_bad_shapes():
    from deeplip_b200 import ops
    x = torch.zeros(1, 4, 4, 12, device='cuda', dtype=torch.bfloat16)          # ldx not a multiple of 8
    w = torch.zeros(8, 64, device='cuda', dtype=torch.bfloat16)
    v = torch.zeros(8, device='cuda')
    with pytest.raises(RuntimeError, match='ldx'):
        ops.conv_igemm(x, w, 12, 8, scale=v, shift=v, slope=v)


@pytest.mark.parametrize('kw', [dict(B=1, T=3, H=32, W=32), dict(B=2, T=6), dict(B=2, T=5, u8=True),
                                dict(B=1, T=1, H=16, W=64), dict(B=1, T=2, H=64, W=88), dict(B=1, T=1, H=24, W=88)])
def test_stem(kw):
    G.stem_case(**kw)


def test_stem_second_generation_ragged_lengths_stacked_rows():
    G.stem_ragged_stacked_case()


@pytest.mark.parametrize('signs', ['mixed_scale', 'neg_slope'])
def test_stem_second_generation_epilogue_modes(signs):
    # the three epilogue modes of stem2_conv3d.cuh (pool-first max, pool-first max + min, general) inside one launch
    G.stem_case(B=2, T=5, u8=True, signs=signs)


def test_stem_second_generation_full_clip():
    # 75 frames: 38 frame pairs per clip, the last one a half pair; more work units than one wave of CTAs would need
    G.stem_case(B=3, T=75, u8=True)


@pytest.mark.parametrize('kw', [dict(), dict(B=2, T=50, C=512, lengths=[50, 7]), dict(B=1, T=3, C=24)])
def test_stat_pool(kw):
    G.stat_pool_case(**kw)
    G.stat_pool_case(B=3, T=100, C=520, lengths=[100, 37, 2])
    G.stat_pool_case(B=1, T=2, C=8)


def test_attn_stat_pool():
    G.attn_pool_case()


@pytest.mark.parametrize('kw', [dict(), dict(HW=4, C=128), dict(B=2, T=40, HW=36, C=64)])
def test_frame_pool_temporal_mean(kw):
    G.frame_pool_case(**kw)
    G.frame_pool_case(B=1, T=1)


def test_fusion_kernels():
    G.fusion_case()


def test_scoring_kernels():
    G.scoring_case()
    G.scoring_case(n_utt=10, D=6, n_trials=1)


def test_scoring_empty_list():
    from deeplip_b200 import ops
    emb = torch.randn(4, 16, device='cuda')
    e = torch.zeros(0, dtype=torch.int32, device='cuda')
    assert ops.cosine_score_trials(emb, e, e).numel() == 0


@pytest.mark.parametrize('kw', [dict(), dict(feat_type='logfbank', n_feat=60), dict(feat_type='fbank', n_feat=24),
                                dict(B=3, nsamp=20000, lengths=[20000, 12345, 300]), dict(B=1, nsamp=48000),
                                dict(gen=1), dict(gen=1, B=3, nsamp=20000, lengths=[20000, 12345, 300]),
                                dict(feat_type='stft'), dict(feat_type='stft', stft_pad='constant', B=2, nsamp=48000),
                                dict(feat_type='stft', B=3, nsamp=20000, lengths=[20000, 12345, 700]),
                                dict(B=2, nsamp=160000, lengths=[160000, 51234]),
                                dict(feat_type='logfbank', n_feat=60, B=2, nsamp=30000, lengths=[401, 30000]),
                                dict(delta=True, B=3, nsamp=20000, lengths=[20000, 12345, 300]),
                                dict(delta=2, feat_type='logfbank', n_feat=60), dict(delta=1, n_feat=13),
                                dict(delta=True, feat_type='stft', B=2, nsamp=8000)])
def test_frontend(kw):
    G.frontend_case(**kw)


def test_avgpool_in_the_conv_epilogue_is_bit_identical_to_the_pooling_kernel():
    G.avgpool_fused_case()
