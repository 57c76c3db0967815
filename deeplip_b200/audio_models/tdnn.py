"""Drop-in for models/audio_models/tdnn.py: TDNN_Block (:7-43) and SpeakerEmbNet (:45-111).

Same constructor (`opts = model_opts` dict, reads opts[opts['arch']]), same parameter names
(tdnn.{i}.context_layer / .bn, pooling.*, fc1, bn1, fc2, bn2) and same methods:
    extract_embedding(x: (B,F,T) f32) -> (xv, x_a)      (tdnn.py:89-101)
    forward(x) -> act(bn2(xv))                           (tdnn.py:103-111)
Every dilated Conv1d + BN + LeakyReLU(0.2) and both Linear heads run as tcgen05 implicit GEMMs
(conv1d == conv2d with H = R = 1) on channels-last bf16 activations; pooling is a fused reduction.
Inference (eval) only; ragged batches pass `lengths` (valid input frames per utterance).
"""
import torch
import torch.nn as nn

from .. import ops, packing
from .pooling import MeanStdPooling, AttentiveStatPooling

LRELU = 0.2


class TDNN_Block(nn.Module):
    def __init__(self, input_dim, output_dim, dilation, padding=0, stride=1, bn_first=True):
        super().__init__()
        kernel_size = len(dilation)
        dilation = (dilation[-1] - dilation[0]) // (len(dilation) - 1) if len(dilation) > 1 else 1
        if padding != 0 or stride != 1:
            raise NotImplementedError('SpeakerEmbNet only builds valid, stride-1 TDNN blocks (tdnn.py:60)')
        self.context_layer = nn.Conv1d(input_dim, output_dim, kernel_size=kernel_size, stride=stride,
                                       padding=padding, dilation=dilation)
        self.bn = nn.BatchNorm1d(output_dim)
        self.activation = nn.LeakyReLU(negative_slope=LRELU)
        self.bn_first = bn_first
        if not bn_first:
            raise NotImplementedError('bn_first=False (act before BN) is not configured anywhere in the reference '
                                      '(conf/*.yaml: bn_first: True) and is not built')
        self.k, self.d = kernel_size, dilation
        self.cin, self.cout = input_dim, output_dim
        self.cout_pad = packing.ceil_to(output_dim, 8)
        self._pk = None

    def _packed(self):
        if self._pk is None:
            cl, bn = self.context_layer, self.bn
            s, h = packing.fold_bn(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
                                   conv_bias=cl.bias.detach(), eps=bn.eps)
            self._pk = dict(w=packing.pack_conv1d_weight(cl.weight.detach(), self.cout_pad),
                            s=packing.pad_vec(s, self.cout_pad), h=packing.pad_vec(h, self.cout_pad),
                            a=torch.full((self.cout_pad,), LRELU, device=cl.weight.device))
        return self._pk

    def forward_ntc(self, x):
        """x: (B,T,ld) bf16 -> (B,T-(k-1)d,cout_pad) bf16."""
        pk = self._packed()
        B, T, ld = x.shape
        cin = ld if ld == packing.ceil_to(self.cin, 64) else self.cin     # zero-padded channels are real zeros
        y, _ = ops.conv_igemm(x.view(B, 1, T, ld), pk['w'], cin, self.cout_pad, 1, self.k, (1, 1), (0, 0),
                              (1, self.d), pk['s'], pk['h'], pk['a'])
        return y.view(B, y.shape[2], self.cout_pad)

    def forward(self, x):
        """Reference signature (B,C,T) f32 -> (B,C',T') f32."""
        y = self.forward_ntc(ops.nct_to_ntc_bf16(x, ld=packing.ceil_to(x.shape[1], 8)))
        return y[:, :, :self.cout].permute(0, 2, 1).float()


class SpeakerEmbNet(nn.Module):
    def __init__(self, opts):
        super().__init__()
        opts = opts[opts['arch']]
        context = opts['context']
        input_dim = opts['input_dim']
        hidden_dim = opts['hidden_dim']
        layers_num = opts['tdnn_layers']
        embedding_dim = opts['embedding_dim']
        attention_hidden_size = opts['attention_hidden_size']
        self.bn_first = opts['bn_first']
        self.activation = nn.LeakyReLU(negative_slope=LRELU)
        self.input_dim = input_dim
        layers = []
        for i in range(layers_num):
            layers.append(TDNN_Block(input_dim, hidden_dim[i], dilation=context[i], stride=1, bn_first=self.bn_first))
            input_dim = hidden_dim[i]
        self.tdnn = nn.Sequential(*layers)
        self.trunk_out = hidden_dim[-1]
        self.context_loss = sum((b.k - 1) * b.d for b in layers)      # frames eaten by the valid convs
        if opts['pooling'] == 'statistic':
            self.pooling = MeanStdPooling()
        elif opts['pooling'] == 'attentive_statistic':
            self.pooling = AttentiveStatPooling(hidden_dim[-1], attention_hidden_size)
        elif opts['pooling'] in ('average', 'mono_head_attention'):
            # 'average' raises a shape error in the reference itself (tdnn.py:69,91-93);
            # MonoHeadAttention is CUDA-ctor-only and never configured.
            raise NotImplementedError("pooling '%s' is not runnable in the reference and is not built" % opts['pooling'])
        else:
            raise NotImplementedError('Other pooling method has not implemented.')
        self.fc1 = nn.Linear(hidden_dim[-1] * 2, embedding_dim)
        self.bn1 = nn.BatchNorm1d(embedding_dim)
        self.fc2 = nn.Linear(embedding_dim, embedding_dim)
        self.bn2 = nn.BatchNorm1d(embedding_dim)
        self.embedding_dim = embedding_dim
        self._pk = None

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = {(k[7:] if k.startswith('module.') else k): v for k, v in state_dict.items()}
        out = super().load_state_dict(sd, strict=strict, **kw)
        self.invalidate()
        return out

    def invalidate(self):
        self._pk = None
        for m in self.modules():
            if isinstance(m, (TDNN_Block, AttentiveStatPooling)):
                m._pk = None

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self.invalidate()
        return out

    def _packed(self):
        if self._pk is None:
            E = self.embedding_dim
            dev = self.fc1.weight.device
            s1, h1 = packing.fold_bn(self.bn1.weight.detach(), self.bn1.bias.detach(), self.bn1.running_mean,
                                     self.bn1.running_var, conv_bias=self.fc1.bias.detach(), eps=self.bn1.eps)
            s2, h2 = packing.fold_bn(self.bn2.weight.detach(), self.bn2.bias.detach(), self.bn2.running_mean,
                                     self.bn2.running_var, eps=self.bn2.eps)
            self._pk = dict(w1=packing.pack_linear_weight(self.fc1.weight.detach()), s1=s1, h1=h1,
                            b1=self.fc1.bias.detach().float().contiguous(),
                            w2=packing.pack_linear_weight(self.fc2.weight.detach()),
                            b2=self.fc2.bias.detach().float().contiguous(), s2=s2, h2=h2,
                            one=torch.ones(E, device=dev), lrelu=torch.full((E,), LRELU, device=dev))
        return self._pk

    @property
    def min_frames(self):
        """Fewest feature frames an utterance needs: two must survive the valid convs (unbiased std)."""
        return self.context_loss + 2

    def embed_ntc(self, x_ntc, lengths=None):
        """x_ntc: (B,T,ld) bf16 channels-last features; lengths: valid *input* frames (int32, CUDA)."""
        if self.training:
            raise RuntimeError('deeplip_b200.SpeakerEmbNet is inference-only: call .eval()')
        pk = self._packed()
        if x_ntc.shape[1] - self.context_loss < 2:
            # the reference's Conv1d raises on an input shorter than its kernel, and one surviving frame has no
            # unbiased std; fail here instead of returning NaN embeddings
            raise ValueError('utterance of %d feature frames is shorter than the TDNN receptive field (%d frames '
                             'needed)' % (x_ntc.shape[1], self.min_frames))
        for blk in self.tdnn:
            x_ntc = blk.forward_ntc(x_ntc)
        if lengths is not None:
            # Rows of a ragged batch shorter than `min_frames` cannot be detected here without a device sync: their
            # pooled std is NaN (one frame) and so is their embedding.  Callers that own host lengths check them
            # first (HostPipeline.run does; `min_frames` is the bound).
            lengths = (lengths - self.context_loss).clamp_(min=1).to(torch.int32)
        _, pooled = self.pooling.pool_ntc(x_ntc, self.trunk_out, lengths)          # (B, 2C) bf16
        B = pooled.shape[0]
        E = self.embedding_dim
        # fc1: x_a = W1 p + b1 (f32 side output) ; h = lrelu(bn1(x_a)) (bf16 main output)
        h, x_a = ops.conv_igemm(pooled.view(B, 1, 1, -1), pk['w1'], pooled.shape[1], E, scale=pk['s1'],
                                shift=pk['h1'], slope=pk['lrelu'], want_f32=True, scale2=pk['one'], shift2=pk['b1'])
        _, xv = ops.conv_igemm(h.view(B, 1, 1, E), pk['w2'], E, E, want_bf16=False, want_f32=True,
                               scale2=pk['one'], shift2=pk['b2'])
        return xv, x_a

    def extract_embedding(self, x, lengths=None):
        return self.embed_ntc(ops.nct_to_ntc_bf16(x, ld=packing.ceil_to(x.shape[1], 64)), lengths)

    def forward(self, x):
        xv, _ = self.extract_embedding(x)
        pk = self._packed()
        _, y = ops.affine_act(xv, pk['s2'], pk['h2'], LRELU, want_bf16=False, want_f32=True)
        return y
