"""compute-sanitizer target: the second-generation stem (two shapes) and the conv with K4 in its epilogue, at sizes the
tool finishes in seconds.   compute-sanitizer --tool memcheck python tools/sanitize_stem2.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import gpu_checks as G
print(G.stem_case(B=1, T=2, H=24, W=88))
print(G.stem_case(B=1, T=3, u8=True))
print(G.stem_ragged_stacked_case(B=2, T=5))
print(G.avgpool_fused_case(N=1875))          # 16 875 rows: just past the pair kernel's threshold of 128 m blocks
print('sanitizer run done')
