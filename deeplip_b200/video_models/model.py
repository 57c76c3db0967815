"""Drop-in for models/video_models/model.py: Lipreading (:61-105) with extract_feats=True.

forward(x: (B,1,T,H,W) f32, lengths) -> (B,T,512) f32, exactly the reference contract used by
train_fusion.py:71-76, 348, 400 and train_video.py:99-106.  The Conv3d stem runs as one fused
tcgen05 kernel (conv + BN + PReLU + max-pool, channels-last output), the ResNet-18 trunk as
implicit-GEMM kernels, the pooling tail as a warp-shuffle reduction.
"""
import torch
import torch.nn as nn

from .. import ops, packing
from .resnet import ResNet, BasicBlock
from .tcn import MultiscaleMultibranchTCN


def threeD_to_2D_tensor(x):
    n_batch, n_channels, s_time, sx, sy = x.shape
    return x.transpose(1, 2).reshape(n_batch * s_time, n_channels, sx, sy)


class Lipreading(nn.Module):
    def __init__(self, hidden_dim=256, backbone_type='resnet', num_classes=500, relu_type='prelu',
                 tcn_options={}, width_mult=1.0, extract_feats=False):
        super().__init__()
        self.extract_feats = extract_feats
        self.backbone_type = backbone_type
        if backbone_type != 'resnet':
            # conf/video_config.json:2 and conf/fusion_config.yaml:75 both select 'resnet'
            raise NotImplementedError("only backbone_type='resnet' is on the B200 hot path (SURVEY 2)")
        self.tcn = None
        if not extract_feats:
            if len(tcn_options['kernel_size']) == 1:
                raise NotImplementedError('the single-branch TCN is not configured by the reference '
                                          '(tcn_kernel_size: [3, 5, 7]); only the multi-scale head is built')
            self.tcn = MultiscaleMultibranchTCN(
                input_size=512,
                num_channels=[hidden_dim * len(tcn_options['kernel_size']) * tcn_options['width_mult']] * tcn_options['num_layers'],
                num_classes=num_classes, tcn_options=tcn_options, dropout=tcn_options['dropout'],
                relu_type=relu_type, dwpw=tcn_options['dwpw'])
        self.frontend_nout = 64
        self.backend_out = 512
        self.trunk = ResNet(BasicBlock, [2, 2, 2, 2], relu_type=relu_type)
        frontend_relu = nn.PReLU(num_parameters=self.frontend_nout) if relu_type == 'prelu' else nn.ReLU()
        self.frontend3D = nn.Sequential(
            nn.Conv3d(1, self.frontend_nout, kernel_size=(5, 7, 7), stride=(1, 2, 2), padding=(2, 3, 3), bias=False),
            nn.BatchNorm3d(self.frontend_nout),
            frontend_relu,
            nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1)))
        self._pk = None

    # -- checkpoints: tolerate 'module.' prefixes and the TCN head keys a reference checkpoint carries
    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = {}
        for k, v in state_dict.items():
            k = k[7:] if k.startswith('module.') else k
            if k.startswith('tcn.') and self.tcn is None:
                continue
            sd[k] = v
        out = super().load_state_dict(sd, strict=strict, **kw)
        self.invalidate()
        return out

    def invalidate(self):
        self._pk = None
        self.trunk.invalidate()
        if getattr(self, 'tcn', None) is not None:
            self.tcn.invalidate()

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self.invalidate()
        return out

    def _packed(self):
        if self._pk is None:
            conv, bn, act = self.frontend3D[0], self.frontend3D[1], self.frontend3D[2]
            pk = {'w': packing.pack_stem_weight(conv.weight.detach())}
            pk['s'], pk['h'] = packing.fold_bn(bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                                               bn.running_var, eps=bn.eps)
            if isinstance(act, nn.PReLU):
                pk['a'] = act.weight.detach().float().expand(64).contiguous()
            else:
                pk['a'] = torch.zeros(64, device=conv.weight.device)
            self._pk = pk
        return self._pk

    def stem_prepass(self, x, lengths=None):
        """The stem's pre-pass alone (x -> the cached workspace), for callers that overlap it with other work on another
        stream and then call trunk_maps / utterance_embedding with prepassed=True on the same x."""
        pk = self._packed()
        ops.stem_conv3d(x, None, None, None, None, lengths=lengths, phase='prepass')

    def trunk_maps(self, x, lengths=None, avgpool=False, prepassed=False):
        """x: (B,T,H,W) f32 normalised frames or (B,T,Hraw,Wraw) uint8 raw crops -> (B*T,3,3,512) bf16; with
        avgpool=True the per-frame pooled features (B*T,512) f32 instead (K4's spatial half in the last conv's epilogue).
        lengths (int32 CUDA, optional): frames at or beyond a clip's length enter the stem as zero normalised
        frames (what pad_packed_collate produces), whatever the padding bytes of a raw u8 batch are."""
        if self.training:
            raise RuntimeError('deeplip_b200.Lipreading is inference-only: call .eval() (BN uses running stats)')
        pk = self._packed()
        B, T = x.shape[0], x.shape[1]
        H, W = (88, 88) if x.dtype == torch.uint8 else (x.shape[2], x.shape[3])
        if self.trunk.halo_enabled(W // 4):
            # stem writes straight into the stacked-rows layout layer1's halo kernel consumes
            buf = self.trunk.stacked_buffers(B * T, H // 4, W // 4, x.device, 2 * len(self.trunk.layer1) + 1)[-1]
            ops.stem_conv3d(x, pk['w'], pk['s'], pk['h'], pk['a'], out=buf, lengths=lengths,
                            phase='main' if prepassed else 'both')
            return self.trunk.forward_nhwc(buf, stacked_H=H // 4, avgpool=avgpool)
        y = ops.stem_conv3d(x, pk['w'], pk['s'], pk['h'], pk['a'], lengths=lengths, phase='main' if prepassed else 'both')
        return self.trunk.forward_nhwc(y, avgpool=avgpool)

    def forward(self, x, lengths=None):
        B, C, T, H, W = x.size()
        assert C == 1
        feats = self.trunk_maps(x[:, 0], avgpool=True).view(B, T, -1)
        if self.extract_feats:
            return feats                  # (B, T, 512); lengths unused when extract_feats (model.py:105)
        return self.tcn(feats, lengths, B)

    def utterance_embedding(self, x, lengths=None, prepassed=False):
        """Fused form of train_fusion.py:400: mean over the (valid) frames of each clip -> (B,512).
        x as in trunk_maps; lengths: int32 CUDA tensor of valid frame counts (zero-padded tails)."""
        B, T = x.shape[0], x.shape[1]
        feats = self.trunk_maps(x, lengths, avgpool=True, prepassed=prepassed)          # (B*T, 512) f32
        return ops.temporal_mean(feats, B, T, lengths=lengths)
