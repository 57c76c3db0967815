"""Run every kernel-level check, never stop at the first failure, write gpurun_out/diag.json."""
import json, os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import gpu_checks as G

only = sys.argv[1:]
res = {}
def run(name, fn, **kw):
    if only and not any(o in name for o in only):
        return
    t0 = time.time()
    try:
        res[name] = {'ok': True, 'metrics': fn(**kw)}
    except AssertionError as e:
        res[name] = {'ok': False, 'assert': str(e)[:400]}
    except Exception as e:
        res[name] = {'ok': False, 'error': repr(e)[:400], 'tb': traceback.format_exc()[-600:]}
        try:
            torch.cuda.synchronize()
        except Exception as e2:
            res[name]['sync'] = repr(e2)[:200]
    res[name]['s'] = round(time.time() - t0, 2)
    print(name, json.dumps(res[name]), flush=True)

print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
run('fusion', G.fusion_case)
run('scoring', G.scoring_case)
run('stat_pool', G.stat_pool_case)
run('stat_pool_ragged', G.stat_pool_case, B=3, T=100, C=520, lengths=[100, 37, 2])
run('frame_pool', G.frame_pool_case)
run('frontend_mfcc', G.frontend_case)
run('frontend_logfbank60', G.frontend_case, feat_type='logfbank', n_feat=60)
run('frontend_ragged', G.frontend_case, B=3, nsamp=20000, lengths=[20000, 12345, 300])
for k, kw in G.CONV_CASES.items():
    run('conv_' + k, G.conv_case, **kw)
run('halo_22', G.halo_case)
run('halo_8x40_nores', G.halo_case, N=3, H=8, W=40, residual=False)
run('halo_big', G.halo_case, N=300, H=22, W=22)
run('attn_pool', G.attn_pool_case)
run('stem_f32_small', G.stem_case, B=1, T=3, H=32, W=32)
run('stem_f32', G.stem_case)
run('stem_u8', G.stem_case, u8=True)
run('video_model', G.video_model_case)
run('video_golden', G.video_golden_case)
run('video_tcn', G.video_tcn_case)
run('audio_model_etdnn', G.audio_model_case)
run('audio_model_tdnn_attn', G.audio_model_case, arch='tdnn', pooling='attentive_statistic')
run('audio_golden', G.audio_golden_case)
run('audio_resnet_avg', G.audio_resnet_case)
run('audio_resnet_stat', G.audio_resnet_case, pooling='statistic')
run('fusion_golden', G.fusion_golden_case)
import tempfile
run('scoring_full_grid', G.scoring_full_case, tmpdir=tempfile.mkdtemp(), kind='grid')
run('scoring_full_lomgrid', G.scoring_full_case, tmpdir=tempfile.mkdtemp(), kind='lomgrid')
run('pipeline', G.pipeline_case)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'diag.json'), 'w'), indent=1)
print('PASS' if all(r['ok'] for r in res.values()) else 'FAIL', sum(r['ok'] for r in res.values()), '/', len(res))
