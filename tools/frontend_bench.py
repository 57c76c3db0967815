import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplip_b200 import ops, synth
from lin_bench import timeit
wav = torch.from_numpy(synth.speech_like_audio(list(range(64)), nsamp=48000, seed=1)).cuda()
t = timeit(lambda: ops.frontend_features(wav, 'mfcc', 24, True), n=20)
print('frontend (frames + cmvn), 64 x 3 s: %.1f us' % t)
