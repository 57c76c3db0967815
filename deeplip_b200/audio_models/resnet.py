"""Audio "ResNet" SpeakerEmbNet for `arch: resnet` (conf/audio_config.yaml:93-102) -- BUILD-DEFINED.

The reference imports `models.resnet.SpeakerEmbNet` (train_audio.py:64-66) but ships no such file
(SURVEY D1): only its hyper-parameters (input_dim 1, hidden_dim [64,128,256], residual_block_layers
[3,3,3], fc_layers 1, embedding_dim 256, pooling average) and its call convention survive --
input `feat.unsqueeze(1)` = (B,1,F,T) (train_audio.py:183-184) and `extract_embedding(x) -> (xv, x_a)`
(:250-252).  There is therefore NO reference behaviour to be identical to; parity for this class is
against the fp32 restatement of this very definition (oracle/models_ref.py: audio_resnet_*), and is
"unpinned" with respect to DeepLip.  The definition reuses the only 2-D residual block the reference
does ship (models/video_models/resnet.py:28-69, ReLU flavour):

    conv3x3(1->C0) + BN + ReLU
    stage i: residual_block_layers[i] x BasicBlock(hidden_dim[i]), stride 2 (+ 1x1-s2 conv/BN skip) from i >= 1
    pooling: 'average' = mean over (F', T') per channel;  'statistic' = mean || unbiased std over (F', T')
    fc1: Linear(C_last [x2] -> embedding_dim)  -> x_a ; fc_layers == 1, so xv = x_a ; forward() = ReLU(bn1(xv))

All convolutions run on the tcgen05 implicit-GEMM kernel; the single input channel is carried as an
8-channel zero-padded bf16 tensor (16 B per pixel, the TMA minimum).  Batches must be equal-length (zero
padding is not exact through padded 2-D convolutions); the reference loops run B=1 anyway.
"""
import torch
import torch.nn as nn

from .. import ops, packing
from ..video_models.resnet import BasicBlock, downsample_basic_block


class SpeakerEmbNet(nn.Module):
    def __init__(self, opts):
        super().__init__()
        o = opts[opts['arch']] if 'arch' in opts else opts
        if o.get('input_dim', 1) != 1:
            raise NotImplementedError('audio resnet takes one input channel (feat.unsqueeze(1))')
        hidden = list(o['hidden_dim'])
        blocks = list(o['residual_block_layers'])
        if o.get('fc_layers', 1) != 1:
            raise NotImplementedError('fc_layers != 1 is not configured by the reference')
        self.pooling_type = o.get('pooling', 'average')
        if self.pooling_type not in ('average', 'statistic'):
            raise NotImplementedError('Other pooling method has not implemented.')
        self.conv1 = nn.Conv2d(1, hidden[0], kernel_size=3, stride=1, padding=1, bias=False)
        self.bn0 = nn.BatchNorm2d(hidden[0])
        inplanes = hidden[0]
        for i, (planes, nb) in enumerate(zip(hidden, blocks)):
            layers = []
            for b in range(nb):
                stride = 2 if (i > 0 and b == 0) else 1
                ds = downsample_basic_block(inplanes, planes, stride) if (stride != 1 or inplanes != planes) else None
                layers.append(BasicBlock(inplanes, planes, stride, ds, relu_type='relu'))
                inplanes = planes
            setattr(self, 'layer%d' % (i + 1), nn.Sequential(*layers))
        self.num_stages = len(hidden)
        self.out_channels = inplanes
        pooled = inplanes * (2 if self.pooling_type == 'statistic' else 1)
        self.embedding_dim = o['embedding_dim']
        self.fc1 = nn.Linear(pooled, self.embedding_dim)
        self.bn1 = nn.BatchNorm1d(self.embedding_dim)
        self._pk = None

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = {(k[7:] if k.startswith('module.') else k): v for k, v in state_dict.items()}
        out = super().load_state_dict(sd, strict=strict, **kw)
        self.invalidate()
        return out

    def invalidate(self):
        self._pk = None
        for m in self.modules():
            if isinstance(m, BasicBlock):
                m._pk = None

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self.invalidate()
        return out

    def _packed(self):
        if self._pk is None:
            dev = self.conv1.weight.device
            C0 = self.conv1.out_channels
            s0, h0 = packing.fold_bn(self.bn0.weight.detach(), self.bn0.bias.detach(), self.bn0.running_mean,
                                     self.bn0.running_var, eps=self.bn0.eps)
            E = self.embedding_dim
            Ep = packing.ceil_to(E, 8)
            s1, h1 = packing.fold_bn(self.bn1.weight.detach(), self.bn1.bias.detach(), self.bn1.running_mean,
                                     self.bn1.running_var, eps=self.bn1.eps)
            self._pk = dict(w0=packing.pack_conv_weight(self.conv1.weight.detach()), s0=s0, h0=h0,
                            a0=torch.zeros(C0, device=dev),
                            wf=packing.pack_linear_weight(self.fc1.weight.detach(), Ep),
                            bf=packing.pad_vec(self.fc1.bias.detach(), Ep), one=torch.ones(Ep, device=dev),
                            s1=s1, h1=h1, Ep=Ep)
        return self._pk

    def extract_embedding(self, x):
        """x: (B,1,F,T) f32 -> (xv, x_a), both (B, embedding_dim) f32."""
        if self.training:
            raise RuntimeError('deeplip_b200 audio SpeakerEmbNet is inference-only: call .eval()')
        pk = self._packed()
        B, C, Fd, T = x.shape
        assert C == 1
        x8 = ops.nct_to_ntc_bf16(x.reshape(B * Fd, 1, T), ld=8).view(B, Fd, T, 8)      # 1 channel + 7 zero lanes
        h, _ = ops.conv_igemm(x8, pk['w0'], 8, self.conv1.out_channels, 3, 3, (1, 1), (1, 1), (1, 1),
                              pk['s0'], pk['h0'], pk['a0'])
        for i in range(self.num_stages):
            for blk in getattr(self, 'layer%d' % (i + 1)):
                h = blk.forward_nhwc(h)
        N, P, Q, Cc = h.shape
        if self.pooling_type == 'average':
            _, pooled = ops.frame_pool_temporal_mean(h, B, 1, want_frames=False, want_mean=True)      # (B, C)
        else:
            pooled, _ = ops.stat_pool(h.view(B, P * Q, Cc), Cc, want_bf16=False)                      # (B, 2C)
        pb, _ = ops.affine_act(pooled, ld=packing.ceil_to(pooled.shape[1], 8))
        _, xa = ops.conv_igemm(pb.view(B, 1, 1, -1), pk['wf'], pooled.shape[1], pk['Ep'], want_bf16=False,
                               want_f32=True, scale2=pk['one'], shift2=pk['bf'])
        xa = xa[:, :self.embedding_dim].contiguous()
        return xa, xa

    def forward(self, x):
        xv, _ = self.extract_embedding(x)
        pk = self._packed()
        _, y = ops.affine_act(xv, pk['s1'], pk['h1'], 0.0, want_bf16=False, want_f32=True)
        return y
