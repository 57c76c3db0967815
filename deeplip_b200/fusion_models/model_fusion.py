"""Drop-in for models/fusion_models/model_fusion.py: Linearfusion (:10-24), model_fusion() (:26-28),
plus the fusion the reference actually executes at test time -- Trainer.feature_normalize + concat
(train_fusion.py:233-238, 405-410)."""
import torch
import torch.nn as nn

from .. import ops, packing

LRELU = 0.2


def feature_normalize(data):
    """train_fusion.py:233-238 for one modality: row z-norm, unbiased std, no eps.  (B,D) f32."""
    z = ops.znorm_concat(data, data)            # both halves identical; keep the first
    return z[:, :data.shape[1]].contiguous()


def concat_fusion(xv_audio, em_video, l2norm=False):
    """cat([znorm(audio), znorm(video)], 1) in one kernel (train_fusion.py:405-410)."""
    return ops.znorm_concat(xv_audio, em_video, biased=False, video_first=False, l2norm=l2norm)


class Linearfusion(nn.Module):
    def __init__(self, input_size, hidden_size, num_classes, extract_feats):
        super().__init__()
        self.extract_feats = extract_feats
        self.fc1 = nn.Linear(input_size, hidden_size)
        self.bn1 = nn.BatchNorm1d(hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.activation = nn.LeakyReLU(negative_slope=LRELU)
        self.input_size, self.hidden_size = input_size, hidden_size
        self._pk = None

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = {(k[7:] if k.startswith('module.') else k): v for k, v in state_dict.items()}
        out = super().load_state_dict(sd, strict=strict, **kw)
        self._pk = None
        return out

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._pk = None
        return out

    def _packed(self):
        if self._pk is None:
            Hd = self.hidden_size
            dev = self.fc1.weight.device
            s1, h1 = packing.fold_bn(self.bn1.weight.detach(), self.bn1.bias.detach(), self.bn1.running_mean,
                                     self.bn1.running_var, conv_bias=self.fc1.bias.detach(), eps=self.bn1.eps)
            self._pk = dict(w1=packing.pack_linear_weight(self.fc1.weight.detach()), s1=s1, h1=h1,
                            w2=packing.pack_linear_weight(self.fc2.weight.detach()),
                            b2=self.fc2.bias.detach().float().contiguous(),
                            one=torch.ones(Hd, device=dev), lrelu=torch.full((Hd,), LRELU, device=dev))
        return self._pk

    def forward(self, x):
        if self.training:
            raise RuntimeError('deeplip_b200.Linearfusion is inference-only: call .eval()')
        pk = self._packed()
        B = x.shape[0]
        xb, _ = ops.affine_act(x, ld=packing.ceil_to(self.input_size, 8))
        x1b, x1 = ops.conv_igemm(xb.view(B, 1, 1, -1), pk['w1'], self.input_size, self.hidden_size, scale=pk['s1'],
                                 shift=pk['h1'], slope=pk['lrelu'], want_f32=True, scale2=pk['s1'], shift2=pk['h1'],
                                 f32_slope=LRELU)
        if self.extract_feats:
            return x1
        _, out = ops.conv_igemm(x1b.view(B, 1, 1, -1), pk['w2'], self.hidden_size, self.hidden_size,
                                want_bf16=False, want_f32=True, scale2=pk['one'], shift2=pk['b2'])
        return out


def model_fusion(input_size, hidden_size, num_classes, extract_feats):
    return Linearfusion(input_size, hidden_size, num_classes, extract_feats)
