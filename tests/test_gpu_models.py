"""-m gpu: model-level parity of the drop-in modules and the batched AV pipeline vs the CPU oracle and
the committed golden vectors (generated from the reference's own modules, oracle/gen_golden.py).
Tolerances are north_star's: embedding cosine >= 0.999, per-trial score within 1e-3, EER within
0.05 % absolute, bit-exact trial indexing."""
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available(), reason='needs a CUDA device (no CPU fallback)')]

if torch.cuda.is_available():
    import gpu_checks as G
else:
    G = None


def test_video_model_vs_oracle():
    G.video_model_case()


def test_video_fused_layer2_entry_is_bit_identical():
    G.video_fused_entry_case()


def test_video_guarded_layer2_path():
    G.video_guarded_case()


def test_video_tcn_head_vs_oracle_and_reference_golden():
    G.video_tcn_case()


def test_video_model_vs_reference_golden():
    G.video_golden_case()


@pytest.mark.parametrize('arch,pooling', [('etdnn', 'statistic'), ('tdnn', 'statistic'),
                                          ('tdnn', 'attentive_statistic')])
def test_audio_model_vs_oracle(arch, pooling):
    G.audio_model_case(arch=arch, pooling=pooling)


@pytest.mark.parametrize('pooling', ['average', 'statistic'])
def test_audio_resnet_build_defined(pooling):
    G.audio_resnet_case(pooling=pooling)


def test_audio_model_vs_reference_golden():
    G.audio_golden_case()


def test_fusion_vs_reference_golden():
    G.fusion_golden_case()


@pytest.mark.parametrize('kind', ['grid', 'lomgrid', 'grid_real', 'lomgrid_real'])
def test_scoring_full_trial_list(tmp_path, kind):
    G.scoring_full_case(str(tmp_path), kind)


def test_trial_list_job_extracts_into_the_gather_table(tmp_path):
    G.trial_list_job_case(tmp_path)


def test_av_pipeline_and_ragged_batch():
    G.pipeline_case()


def test_graphed_extractor_replays_bit_identically():
    G.graphed_extractor_case()


def test_full_size_batch_256_is_batch_invariant():
    G.full_size_batch_invariance_case()


def test_plda_trial_scoring_vs_oracle():
    G.plda_case()


def test_tcd_timit_shaped_ragged_long_utterances():
    G.tcd_timit_ragged_case()


def test_bitwise_determinism_back_to_back():
    G.determinism_case()


def test_bitwise_determinism_at_pair_kernel_size():
    G.determinism_pair_size_case()


def test_modules_are_inference_only_and_fail_loudly_on_cpu():
    from deeplip_b200 import ops
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.l2_normalize(torch.randn(2, 4))


def test_reference_wire_format_roundtrip(tmp_path):
    """per-utterance (1,D) .npy layout (train_fusion.py:414-417) -> eer_cos_grid signature."""
    import numpy as np
    from deeplip_b200.fusion_models import utils as U
    from deeplip_b200.trials import TrialList
    from deeplip_b200 import synth
    p = G.make_trial_file(str(tmp_path / 'trial.txt'), 'grid', n_target=300, n_non=900)
    tl = TrialList.from_file(p)
    emb = synth.structured_embeddings([synth.speaker_of_utt(u) for u in tl.utts], dim=64, seed=2, within=3.0)
    U.save_embeddings(str(tmp_path / 'exp' / 'run1' / 'test_em_grid'), tl.utts, emb)
    eer, thr = U.eer_cos_grid('run1', trial_path=p, root=str(tmp_path / 'exp'))
    eer2, _ = U.eer_cos(tl, emb)
    assert abs(eer - eer2) < 1e-12 and 0 < eer < 0.5


def test_video_embedding_with_fused_avgpool_equals_unfused_and_alone():
    G.avgpool_model_case()


def test_stem_prepass_on_a_side_stream_gives_the_same_bits():
    G.prepass_overlap_case()
