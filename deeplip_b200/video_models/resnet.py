"""Drop-in for models/video_models/resnet.py (BasicBlock :28-69, ResNet :72-127).

Parameters live in ordinary nn.Conv2d / nn.BatchNorm2d / nn.PReLU containers with the reference's
attribute names, so reference checkpoints load unchanged; the containers are never *called* --
forward runs dl_conv_igemm_bf16 (tcgen05 implicit GEMM, BN/PReLU/residual fused) on channels-last
bf16 activations.  Inference (eval) only.
"""
import math
import os
import torch
import torch.nn as nn

from .. import ops, packing

# layer1 (64 -> 64, 3x3, stride 1) runs on the halo-reuse kernel (dl_conv3x3_c64_halo_bf16) unless disabled
USE_HALO = os.environ.get('DL_USE_HALO', '1') != '0'
# Optional: layer2 (128 -> 128, 3x3, stride 1) keeps its activations in the guarded layout (one zero row / column
# after every image) so that the CTA-pair kernel can share one tiled-TMA box of operand A across the three
# horizontal taps.  Off by default: since the tensor-issue fix it runs level with the im2col kernel
# (4.13 vs 4.12 ms per step) while doing 19 % more MMAs.
USE_GUARDED = os.environ.get('DL_USE_GUARDED', '0') != '0'
# layer2's entry block: its 3x3 stride-2 conv and its 1x1 stride-2 downsample conv read the same pixels (the 1x1 window
# is the centre tap of the 3x3 one), so they run as ONE 64 -> 256 conv (channels [0,128) conv1, [128,256) the skip with
# zero weights off the centre tap): the 297 MB layer1 map is read once instead of twice and the tiles are 256 wide.
# Bit-identical results (the extra products are exact zeros).
FUSE_L2_ENTRY = os.environ.get('DL_FUSE_L2_ENTRY', '1') != '0'


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


def downsample_basic_block(inplanes, outplanes, stride):
    return nn.Sequential(nn.Conv2d(inplanes, outplanes, kernel_size=1, stride=stride, bias=False),
                         nn.BatchNorm2d(outplanes))


def _slope_of(act, planes, device):
    if isinstance(act, nn.PReLU):
        w = act.weight.detach().float()
        return (w.expand(planes) if w.numel() == 1 else w).contiguous()
    return torch.zeros(planes, device=device)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, relu_type='relu'):
        super().__init__()
        assert relu_type in ['relu', 'prelu']
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = nn.BatchNorm2d(planes)
        if relu_type == 'relu':
            self.relu1 = nn.ReLU(inplace=True)
            self.relu2 = nn.ReLU(inplace=True)
        else:
            self.relu1 = nn.PReLU(num_parameters=planes)
            self.relu2 = nn.PReLU(num_parameters=planes)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride
        self.inplanes, self.planes = inplanes, planes
        self._pk = None

    def _packed(self):
        if self._pk is None:
            dev = self.conv1.weight.device
            pk = {}
            pk['w1'] = packing.pack_conv_weight(self.conv1.weight.detach())
            pk['s1'], pk['h1'] = packing.fold_bn(self.bn1.weight.detach(), self.bn1.bias.detach(),
                                                 self.bn1.running_mean, self.bn1.running_var, eps=self.bn1.eps)
            pk['a1'] = _slope_of(self.relu1, self.planes, dev)
            pk['w2'] = packing.pack_conv_weight(self.conv2.weight.detach())
            pk['s2'], pk['h2'] = packing.fold_bn(self.bn2.weight.detach(), self.bn2.bias.detach(),
                                                 self.bn2.running_mean, self.bn2.running_var, eps=self.bn2.eps)
            pk['a2'] = _slope_of(self.relu2, self.planes, dev)
            if self.downsample is not None:
                dconv, dbn = self.downsample[0], self.downsample[1]
                pk['wd'] = packing.pack_conv_weight(dconv.weight.detach())
                pk['sd'], pk['hd'] = packing.fold_bn(dbn.weight.detach(), dbn.bias.detach(), dbn.running_mean,
                                                     dbn.running_var, eps=dbn.eps)
                pk['ad'] = torch.ones(self.planes, device=dev)       # identity activation on the skip
                if self.stride == 2 and self.inplanes % 64 == 0:
                    # conv1 and the skip as one conv of 2 x planes outputs: the skip's weights sit on the centre tap
                    kb = (self.inplanes + 63) // 64 * 64
                    wf = torch.zeros((2 * self.planes, 9 * kb), device=dev, dtype=pk['w1'].dtype)
                    wf[:self.planes] = pk['w1']
                    wf[self.planes:, 4 * kb:5 * kb] = pk['wd']
                    pk['wf'] = wf.contiguous()
                    pk['sf'] = torch.cat([pk['s1'], pk['sd']]).contiguous()
                    pk['hf'] = torch.cat([pk['h1'], pk['hd']]).contiguous()
                    pk['af'] = torch.cat([pk['a1'], pk['ad']]).contiguous()
            self._pk = pk
        return self._pk

    @property
    def halo_ok(self):
        return self.stride == 1 and self.inplanes == 64 and self.planes == 64 and self.downsample is None

    def forward_stacked(self, x, H, mid, out):
        """Stacked-rows layout (N, >=H+1, W, 64): both convs on the halo-reuse kernel; mid/out are
        caller-owned buffers whose padding rows are zero."""
        pk = self._packed()
        ops.conv3x3_halo(x, pk['w1'], pk['s1'], pk['h1'], pk['a1'], H, out=mid)
        ops.conv3x3_halo(mid, pk['w2'], pk['s2'], pk['h2'], pk['a2'], H, out=out, residual=x)
        return out

    def forward_guarded(self, x, H, W, bufs, x_guarded):
        """Output (and, when x_guarded, input) in the guarded layout (N, P+1, Q+1, planes) -- one zero row / column
        after every image, see include/deeplip_b200.h.  bufs: three caller-owned zeroed guarded buffers, none of
        them x; returns the one holding the block output.  x not guarded: (N, img_rows >= H, W, C)."""
        pk = self._packed()
        st = (self.stride, self.stride)
        mid, res, out = bufs
        P, Q = mid.shape[1] - 1, mid.shape[2] - 1
        if self.downsample is not None:      # stride-2 entry block: im2col kernel writing the guarded layout
            ops.conv_igemm(x, pk['w1'], self.inplanes, self.planes, 3, 3, st, (1, 1), (1, 1),
                           pk['s1'], pk['h1'], pk['a1'], H=H, W=W, out=mid)
            ops.conv_igemm(x, pk['wd'], self.inplanes, self.planes, 1, 1, st, (0, 0), (1, 1),
                           pk['sd'], pk['hd'], pk['ad'], H=H, W=W, out=res)
        else:
            assert x_guarded and self.stride == 1
            ops.conv_igemm_lin(x, pk['w1'], self.inplanes, self.planes, (P, Q), 3, 3, (1, 1), (1, 1),
                               pk['s1'], pk['h1'], pk['a1'], out=mid)
            res = x
        ops.conv_igemm_lin(mid, pk['w2'], self.planes, self.planes, (P, Q), 3, 3, (1, 1), (1, 1),
                           pk['s2'], pk['h2'], pk['a2'], residual=res, out=out)
        return out

    def forward_fused_entry_split(self, x, H=None, W=None):
        """Entry block of a >= 256-channel layer: conv1 and the skip as one conv whose two halves leave as two dense
        tensors (dl_conv_desc.split_channel); the skip half is declared centre-tap-only, so the CTA-pair kernel
        spends a ninth of a tile on it instead of a separate launch that re-reads the input map."""
        pk = self._packed()
        (mid, res), _ = ops.conv_igemm(x, pk['wf'], self.inplanes, 2 * self.planes, 3, 3, (2, 2), (1, 1), (1, 1),
                                       pk['sf'], pk['hf'], pk['af'], H=H, W=W, split=(self.planes, True))
        out, _ = ops.conv_igemm(mid, pk['w2'], self.planes, self.planes, 3, 3, (1, 1), (1, 1), (1, 1),
                                pk['s2'], pk['h2'], pk['a2'], residual=res)
        return out

    def forward_fused_entry(self, x, H, W, out):
        """Entry block with conv1 and the skip as one conv (see FUSE_L2_ENTRY).  x: (N, img_rows >= H, W, inplanes);
        out: caller-owned (N, P, Q, 2 planes) buffer; returns it with the block output in channels [0, planes)."""
        pk = self._packed()
        both, _ = ops.conv_igemm(x, pk['wf'], self.inplanes, 2 * self.planes, 3, 3, (2, 2), (1, 1), (1, 1),
                                 pk['sf'], pk['hf'], pk['af'], H=H, W=W,             # [conv1 | skip], pitch 2 planes
                                 center_only_from=self.planes)
        ops.conv_igemm(both, pk['w2'], self.planes, self.planes, 3, 3, (1, 1), (1, 1), (1, 1), pk['s2'], pk['h2'],
                       pk['a2'], residual=both, residual_channel_offset=self.planes, out=out)
        return out

    def forward_pitched(self, x, out):
        """Identity-skip block on activations that live in the first `planes` channels of wider buffers (x, out:
        (N, P, Q, ld >= planes), the pitch the fused entry block leaves behind)."""
        pk = self._packed()
        assert self.downsample is None and self.stride == 1 and x.shape == out.shape
        mid, _ = ops.conv_igemm(x, pk['w1'], self.inplanes, self.planes, 3, 3, (1, 1), (1, 1), (1, 1),
                                pk['s1'], pk['h1'], pk['a1'])
        ops.conv_igemm(mid, pk['w2'], self.planes, self.planes, 3, 3, (1, 1), (1, 1), (1, 1), pk['s2'], pk['h2'],
                       pk['a2'], residual=x, out=out)
        return out

    def forward_nhwc(self, x, H=None, W=None, avgpool=False):
        """x: (N,H,W,inplanes) bf16 -> (N,P,Q,planes) bf16  (reference forward :56-69).  With H (W) given, x is in
        a stacked / guarded layout (N, img_rows >= H, img_cols >= W, C); only blocks with a downsample branch
        accept that.  avgpool=True: returns the global average pool of the block's output instead, (N, planes) f32,
        taken in the last conv's epilogue (resnet.py:125-126 fused, K4)."""
        pk = self._packed()
        st = (self.stride, self.stride)
        out, _ = ops.conv_igemm(x, pk['w1'], self.inplanes, self.planes, 3, 3, st, (1, 1), (1, 1),
                                pk['s1'], pk['h1'], pk['a1'], H=H, W=W)
        if self.downsample is not None:
            res, _ = ops.conv_igemm(x, pk['wd'], self.inplanes, self.planes, 1, 1, st, (0, 0), (1, 1),
                                    pk['sd'], pk['hd'], pk['ad'], H=H, W=W)
        else:
            assert H is None and W is None, 'identity residual needs the dense layout'
            res = x
        out, pooled = ops.conv_igemm(out, pk['w2'], self.planes, self.planes, 3, 3, (1, 1), (1, 1), (1, 1),
                                     pk['s2'], pk['h2'], pk['a2'], residual=res, avgpool=avgpool)
        return pooled if avgpool else out

    def forward(self, x):
        y = self.forward_nhwc(x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))
        return y.permute(0, 3, 1, 2).float()


class ResNet(nn.Module):
    def __init__(self, block, layers, num_classes=1000, relu_type='relu', gamma_zero=False,
                 avg_pool_downsample=False):
        self.inplanes = 64
        self.relu_type = relu_type
        self.gamma_zero = gamma_zero
        if avg_pool_downsample:
            raise NotImplementedError('avg_pool_downsample is never enabled by the reference configs')
        self.downsample_block = downsample_basic_block
        super().__init__()
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        for m in self.modules():       # reference default init, resnet.py:88-96
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        if self.gamma_zero:
            for m in self.modules():
                if isinstance(m, BasicBlock):
                    m.bn2.weight.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = self.downsample_block(inplanes=self.inplanes, outplanes=planes * block.expansion,
                                               stride=stride)
        layers = [block(self.inplanes, planes, stride, downsample, relu_type=self.relu_type)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, relu_type=self.relu_type))
        return nn.Sequential(*layers)

    def invalidate(self):
        for m in self.modules():
            if isinstance(m, BasicBlock):
                m._pk = None

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self.invalidate()
        return out

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self.invalidate()
        return out

    def halo_enabled(self, W):
        return USE_HALO and W >= 8 and all(b.halo_ok for b in self.layer1) and self.layer2[0].downsample is not None

    def guarded_enabled(self, N, P, Q):
        """Worth it only where the CTA-pair kernel runs (>= 128 row blocks): the guards add (P+1)(Q+1)/(PQ) work."""
        l2, l3 = self.layer2, self.layer3
        return (USE_GUARDED and N * (P + 1) * (Q + 1) >= 128 * 128 and l2[0].planes == 128 and
                l2[0].downsample is not None and all(b.downsample is None and b.stride == 1 for b in list(l2)[1:]) and
                l3[0].downsample is not None)

    def fused_entry_enabled(self):
        l2 = self.layer2
        return (FUSE_L2_ENTRY and l2[0].downsample is not None and l2[0].stride == 2 and l2[0].inplanes % 64 == 0 and
                all(b.downsample is None and b.stride == 1 for b in list(l2)[1:]) and
                self.layer3[0].downsample is not None)

    def pitched_buffers(self, N, P, Q, ld, device):
        """Two persistent (N, P, Q, ld) activation buffers (only the first ld/2 channels carry block outputs).
        Cached per shape (ops.BufferCache): a new batch shape never frees buffers an earlier shape -- or a CUDA
        graph captured at it -- still uses."""
        return ops.BUFFERS.get(('pitched', id(self), N, P, Q, ld, str(device)), lambda: [
            torch.zeros((N, P, Q, ld), device=device, dtype=torch.bfloat16) for _ in range(2)])

    def guarded_buffers(self, N, P, Q, device):
        """Persistent zeroed (N, P+1, Q+1, 128) buffers; the kernels never write the guard row / column."""
        return ops.BUFFERS.get(('guarded', id(self), N, P, Q, str(device)), lambda: [
            torch.zeros((N, P + 1, Q + 1, 128), device=device, dtype=torch.bfloat16) for _ in range(3)])

    def stacked_buffers(self, N, H, W, device, count):
        """Persistent zero-initialised (N, H+1, W, 64) activation buffers (padding rows are never written)."""
        return ops.BUFFERS.get(('stacked', id(self), N, H, W, str(device), count), lambda: [
            torch.zeros((N, H + 1, W, 64), device=device, dtype=torch.bfloat16) for _ in range(count)])

    def forward_nhwc(self, x, stacked_H=None, avgpool=False):
        """(N,H,W,64) bf16 -- or stacked rows (N,H+1,W,64) with stacked_H=H -- -> (N,P,Q,512) bf16, before the
        global average pool; with avgpool=True the pooled (N,512) f32 features instead (the pool rides in the epilogue
        of layer4's last conv)."""
        last = list(self.layer4)[-1]
        kw = lambda blk: dict(avgpool=True) if (avgpool and blk is last) else {}
        N, rows, W, _ = x.shape
        H = stacked_H or rows
        if self.halo_enabled(W):
            nb = len(self.layer1)
            bufs = self.stacked_buffers(N, H, W, x.device, 2 * nb + 1)
            if stacked_H is None:
                bufs[-1][:, :H].copy_(x)
                x = bufs[-1]
            for i, blk in enumerate(self.layer1):
                x = blk.forward_stacked(x, H, bufs[2 * i], bufs[2 * i + 1])
            P, Q = (H - 1) // 2 + 1, (W - 1) // 2 + 1
            if self.guarded_enabled(N, P, Q):
                # layer2 in the guarded layout; layer3's stride-2 entry block reads it through im2col pitches
                g = self.guarded_buffers(N, P, Q, x.device)
                x = self.layer2[0].forward_guarded(x, H, W, (g[0], g[1], g[2]), x_guarded=False)
                for blk in list(self.layer2)[1:]:
                    free = [b for b in g if b is not x]
                    x = blk.forward_guarded(x, P, Q, (free[0], None, free[1]), x_guarded=True)
                x = self.layer3[0].forward_nhwc(x, H=P, W=Q)
                rest = list(self.layer3)[1:]
            elif self.fused_entry_enabled():
                a, b = self.pitched_buffers(N, P, Q, 2 * self.layer2[0].planes, x.device)
                x = self.layer2[0].forward_fused_entry(x, H, W, a)
                for blk in list(self.layer2)[1:]:
                    a, b = b, a
                    x = blk.forward_pitched(x, a)
                # layer3 / layer4: entry blocks fused with a split output (their input may sit in a wider pitch)
                for layer in (self.layer3, self.layer4):
                    blocks = list(layer)
                    if blocks[0].planes % 256 == 0 and blocks[0].stride == 2 and blocks[0].downsample is not None and \
                            blocks[0].inplanes % 64 == 0:
                        x = blocks[0].forward_fused_entry_split(x)
                    else:
                        x = blocks[0].forward_nhwc(x)
                    for blk in blocks[1:]:
                        x = blk.forward_nhwc(x, **kw(blk))
                return self._pooled(x, avgpool)
            else:
                x = self.layer2[0].forward_nhwc(x, H=H)
                rest = list(self.layer2)[1:] + list(self.layer3)
        else:
            assert stacked_H is None
            for blk in self.layer1:
                x = blk.forward_nhwc(x)
            rest = list(self.layer2) + list(self.layer3)
        for blk in rest + list(self.layer4):
            x = blk.forward_nhwc(x, **kw(blk))
        return self._pooled(x, avgpool)

    @staticmethod
    def _pooled(x, avgpool):
        """The last block pools in its epilogue; a layer4 of ONE block ends on a fused entry block, which does not: pool
        its maps with the stand-alone kernel (same arithmetic)."""
        if avgpool and x.dim() == 4:
            feats, _ = ops.frame_pool_temporal_mean(x, x.shape[0], 1, want_frames=True, want_mean=False)
            return feats.view(x.shape[0], -1)
        return x

    def forward(self, x):
        """Reference signature: (N,64,H,W) f32 -> (N,512) f32 (resnet.py:120-127)."""
        return self.forward_nhwc(x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16), avgpool=True)
