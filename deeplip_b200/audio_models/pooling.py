"""Drop-ins for models/audio_models/pooling.py: MeanStdPooling (:7-26) and
AttentiveStatPooling (:73-107).  Standalone `forward(x: (B,C,T) f32)` keeps the reference
signature; SpeakerEmbNet calls the channels-last entry points directly."""
import torch
import torch.nn as nn

from .. import ops, packing


class MeanStdPooling(nn.Module):
    def pool_ntc(self, x_ntc, C, lengths=None):
        return ops.stat_pool(x_ntc, C, lengths=lengths)

    def forward(self, x):
        f32, _ = ops.stat_pool(ops.nct_to_ntc_bf16(x, ld=packing.ceil_to(x.shape[1], 8)), x.shape[1],
                               want_bf16=False)
        return f32


class AttentiveStatPooling(nn.Module):
    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.hidden_size = hidden_size
        self.input_size = input_size
        self.W = nn.Parameter(torch.Tensor(hidden_size, input_size))
        self.b = nn.Parameter(torch.Tensor(1, hidden_size))
        self.v = nn.Parameter(torch.Tensor(hidden_size, 1))
        self.k = nn.Parameter(torch.Tensor(1, 1))
        for p in self.parameters():
            nn.init.xavier_normal_(p)
        self._pk = None

    def _packed(self):
        if self._pk is None:
            hp = packing.ceil_to(self.hidden_size, 8)
            self._pk = dict(w=packing.pack_linear_weight(self.W.detach(), hp),
                            b=packing.pad_vec(self.b.detach().reshape(-1), hp),
                            one=packing.pad_vec(torch.ones(self.hidden_size, device=self.W.device), hp),
                            v=self.v.detach().reshape(-1).float().contiguous(),
                            k=float(self.k.detach().reshape(-1)[0]), hp=hp)
        return self._pk

    def pool_ntc(self, x_ntc, C, lengths=None):
        """x_ntc: (B,T,ld) bf16.  h = W x + b on tensor cores (f32 side output), e = v.relu(h)+k,
        softmax over time and the weighted mean/std in one reduction kernel (pooling.py:89-107)."""
        pk = self._packed()
        B, T, ld = x_ntc.shape
        _, h = ops.conv_igemm(x_ntc.view(B, 1, T, ld), pk['w'], C, pk['hp'], want_bf16=False, want_f32=True,
                              scale2=pk['one'], shift2=pk['b'])
        e = ops.attn_logits(h[:, :self.hidden_size].contiguous() if pk['hp'] != self.hidden_size else h,
                            pk['v'], pk['k'])
        return ops.stat_pool(x_ntc, C, lengths=lengths, logits=e.view(B, T))

    def forward(self, x):
        f32, _ = self.pool_ntc(ops.nct_to_ntc_bf16(x, ld=packing.ceil_to(x.shape[1], 8)), x.shape[1])
        return f32
