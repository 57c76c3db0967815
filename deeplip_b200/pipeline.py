"""Batched audio-visual embedding extraction: the B200 form of the reference's per-utterance loop
`Trainer.extract_test_xv_grid` / `extract_test_xv_lomgrid` (train_fusion.py:317-420).

Reference order per utterance (train_fusion.py:386-417):
    wav -> MFCC + CMVN (DataLoader worker)        -> model_audio.extract_embedding -> xv_audio (1,512)
    clips -> /255, centre crop, (x-.421)/.165     -> model_video per clip -> mean over frames -> mean over clips
    feature_normalize x2 -> cat([audio, video])   -> np.save
Here a whole batch of utterances goes through the same arithmetic in 35 kernel launches, nothing
leaves the device, and ragged batches carry `lengths` (zero-padded tails are exact for the video
branch; the audio branch masks its pooling).
"""
import torch

from . import ops


class AVExtractor:
    FUSIONS = ('concat', 'concat_np', 'linear', 'lowfer', 'audio', 'video')

    def __init__(self, audio_model, video_model, fusion='concat', fusion_model=None, feat_type='mfcc',
                 n_feat=24, cmvn=True, l2norm=False, delta=False):
        if fusion not in self.FUSIONS:
            raise NotImplementedError('fusion %r (have %s)' % (fusion, ', '.join(self.FUSIONS)))
        self.audio, self.video = audio_model, video_model
        self.fusion, self.fusion_model = fusion, fusion_model
        self.feat_type, self.n_feat, self.cmvn, self.l2norm, self.delta = feat_type, n_feat, cmvn, l2norm, delta
        # extract(): the stem's pre-pass on a side stream under the audio branch.  Off by default: interleaved A/B at
        # B = 64 in the power-capped state, 3.547 vs 3.560 ms per step (tools/prepass_overlap_ab.py) -- the box's power
        # budget, not free issue slots, is what the step runs against; bit-identical either way.
        self.overlap_prepass = False
        self._side = self._side_done = None

    def audio_embedding(self, wav, wav_lengths=None):
        """wav (B,nsamp) f32 -- or int16 PCM, value / 32768 -- CUDA -> xv (B,E) f32 (LMCL convention: 2nd fc output,
        train_fusion.py:390)."""
        _, feat = ops.frontend_features(wav, self.feat_type, self.n_feat, self.cmvn, lengths=wav_lengths, delta=self.delta)
        frames = None
        if wav_lengths is not None:
            if self.feat_type == 'stft':
                frames = (1 + torch.div(wav_lengths, 160, rounding_mode='floor')).to(torch.int32)
            else:
                frames = torch.where(wav_lengths <= 400, torch.ones_like(wav_lengths),
                                     1 + torch.div(wav_lengths - 400 + 159, 160, rounding_mode='floor')).to(torch.int32)
        xv, _ = self.audio.embed_ntc(feat, frames)
        return xv

    def video_embedding(self, video, video_lengths=None, prepassed=False):
        """video (B,T,H,W) f32 / (B,T,96,96) u8, or (B,G,T,..) for G clips per utterance -> (B,512)."""
        if video.dim() == 5:
            B, G = video.shape[:2]
            em = self.video.utterance_embedding(video.reshape(B * G, *video.shape[2:]),
                                                None if video_lengths is None else video_lengths.reshape(-1))
            return em.view(B, G, -1).mean(dim=1)      # mean over clips (train_fusion.py:401)
        return self.video.utterance_embedding(video, video_lengths, prepassed=prepassed)

    def fuse(self, xv_audio, em_video, out=None):
        """out: optional caller-owned (B, D) f32 rows the fused embeddings are written to (the job's all-gather
        table); the default concat fusions write there directly, the others are copied in."""
        if self.fusion == 'concat':          # F0, the default executed at test time
            return ops.znorm_concat(xv_audio, em_video, l2norm=self.l2norm, out=out)
        if self.fusion == 'concat_np':       # F3, models/fusion_models/utils.py:465-471
            return ops.znorm_concat(xv_audio, em_video, biased=True, video_first=True, l2norm=self.l2norm, out=out)
        if self.fusion == 'linear':          # F1 on cat([audio, video])
            y = self.fusion_model(torch.cat([xv_audio, em_video], dim=1))
        elif self.fusion == 'lowfer':        # F2
            y = ops.lowfer(xv_audio, em_video)
        else:
            y = xv_audio if self.fusion == 'audio' else em_video
        if out is not None:
            out.copy_(y)
            return out
        return y

    @property
    def dim(self):
        """Width of one fused embedding row."""
        ea = getattr(self.audio, 'embedding_dim', 512)
        if self.fusion in ('concat', 'concat_np'):
            return ea + 512
        if self.fusion == 'lowfer':
            return 3 * ea
        if self.fusion == 'linear':
            return self.fusion_model.fc1.out_features if getattr(self.fusion_model, 'extract_feats', True) else \
                self.fusion_model.fc2.out_features
        return ea if self.fusion == 'audio' else 512

    @torch.no_grad()
    def extract(self, wav, video, wav_lengths=None, video_lengths=None, out=None):
        if self.overlap_prepass and self.fusion not in ('video', 'audio') and video.dim() == 4:
            # The stem's pre-pass (HBM-bound, no shared memory, 36 us at B = 64) runs on a side stream under the audio
            # branch, whose kernels leave room for it (the persistent TDNN CTAs use one CTA's worth of shared memory
            # and a third of the registers of an SM; the small kernels do not fill the chip).  Same kernels, same bits.
            main = torch.cuda.current_stream(video.device)
            if self._side is None or self._side.device != video.device:
                self._side = torch.cuda.Stream(device=video.device)
                self._side_done = torch.cuda.Event()
            self._side.wait_stream(main)                # everything enqueued so far, the previous step's stem included
            with torch.cuda.stream(self._side):
                self.video.stem_prepass(video, video_lengths)
                self._side_done.record(self._side)
            xv = self.audio_embedding(wav, wav_lengths)
            main.wait_event(self._side_done)
            em = self.video_embedding(video, video_lengths, prepassed=True)
            return self.fuse(xv, em, out=out)
        xv = self.audio_embedding(wav, wav_lengths) if self.fusion != 'video' else None
        em = self.video_embedding(video, video_lengths) if self.fusion != 'audio' else None
        return self.fuse(xv, em, out=out)


class GraphedExtractor:
    """The whole extraction step (35 kernel launches) captured once into a CUDA graph and replayed:
    static input / output buffers, no per-launch host work.  Shapes are fixed at capture time."""

    def __init__(self, extractor, wav_example, video_example, warmup=2):
        self.ex = extractor
        self.wav = torch.empty_like(wav_example)
        self.video = torch.empty_like(video_example)
        self.wav.copy_(wav_example)
        self.video.copy_(video_example)
        side = torch.cuda.Stream(device=self.wav.device)
        side.wait_stream(torch.cuda.current_stream(self.wav.device))
        # The kernels' persistent buffers (stem workspace, zero-padded trunk layouts) are allocated outside the graph
        # pool and their addresses are baked into the graph: hold references for the graph's lifetime, so that eager
        # calls at other shapes (a dataset's tail batch) can neither free nor resize them (ops.BufferCache).
        with ops.BUFFERS.recording() as pinned:
            with torch.cuda.stream(side):      # first calls build packed weights, set attributes, size caches
                for _ in range(warmup):
                    self.ex.extract(self.wav, self.video)
            torch.cuda.current_stream(self.wav.device).wait_stream(side)
            torch.cuda.synchronize(self.wav.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = self.ex.extract(self.wav, self.video)
        self._pinned_buffers = list(pinned)

    @torch.no_grad()
    def extract(self, wav, video):
        """Same contract as AVExtractor.extract for the captured shapes; returns the graph's static output
        buffer (overwritten by the next call)."""
        if wav.shape != self.wav.shape or video.shape != self.video.shape or video.dtype != self.video.dtype:
            raise RuntimeError('GraphedExtractor was captured for %s / %s' % (tuple(self.wav.shape), tuple(self.video.shape)))
        if wav.data_ptr() != self.wav.data_ptr():
            self.wav.copy_(wav, non_blocking=True)
        if video.data_ptr() != self.video.data_ptr():
            self.video.copy_(video, non_blocking=True)
        self.graph.replay()
        return self.out


class HostPipeline:
    """End-to-end extraction from pinned HOST buffers with the H2D copy of batch i+1 overlapped with the
    kernels of batch i (copy stream + events, two device staging slots) and the D2H of the fused
    embeddings issued asynchronously into pinned memory.  This is the call a user with data on the host
    makes; the reference does one blocking `.to(device)` per clip (train_fusion.py:388, 399)."""

    def __init__(self, extractor, device='cuda', slots=2):
        if slots < 2:
            # with one slot the upload of batch k+1 would overwrite the slot batch k's kernels have not read yet
            raise ValueError('HostPipeline needs slots >= 2 (the upload of batch k+1 overlaps the kernels of batch k)')
        self.ex = extractor
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = slots
        self._stage = [None] * slots
        self._ready = [torch.cuda.Event() for _ in range(slots)]
        self._ready_wav = [torch.cuda.Event() for _ in range(slots)]
        self._free = [torch.cuda.Event() for _ in range(slots)]
        self._out_pool = None       # pinned output rows, grown on demand and kept: cudaHostAlloc costs tens of ms

    def _upload(self, k, wav_h, vid_h, wav_len_h=None, vid_len_h=None):
        slot = k % self.slots
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._free[slot])          # kernels that read this slot are done
            st = self._stage[slot]
            if st is None or st[0].shape != wav_h.shape or st[0].dtype != wav_h.dtype or st[1].shape != vid_h.shape or \
                    st[1].dtype != vid_h.dtype:
                st = [torch.empty(wav_h.shape, dtype=wav_h.dtype, device=self.device),
                      torch.empty(vid_h.shape, dtype=vid_h.dtype, device=self.device),
                      torch.empty((wav_h.shape[0],), dtype=torch.int32, device=self.device),
                      torch.empty((wav_h.shape[0],), dtype=torch.int32, device=self.device), False]
                self._stage[slot] = st
            st[4] = wav_len_h is not None
            if st[4]:                                              # ragged batch: the two length vectors ride along
                st[2].copy_(wav_len_h, non_blocking=True)
                st[3].copy_(vid_len_h, non_blocking=True)
            st[0].copy_(wav_h, non_blocking=True)
            self._ready_wav[slot].record(self.copy_stream)         # the audio branch starts while the crops still upload
            st[1].copy_(vid_h, non_blocking=True)
            self._ready[slot].record(self.copy_stream)

    @torch.no_grad()
    def run(self, batches, post=None):
        """post: optional callable applied to every batch's fused embeddings before they go to the host (bench.py: the
        all_gather of the step's rows); a dist.OverlappedGather runs on its own stream, under the next batch's kernels.
        batches: sequence of (wav_host (B,nsamp) f32 pinned, video_host (B,T,H,W) u8/f32 pinned) or, for ragged
        batches (zero-padded tails, pad_packed_collate convention), (wav, video, wav_lengths, video_lengths) with the
        lengths as pinned int32 host vectors.  Returns the list of fused embeddings as pinned host tensors (valid after
        the final synchronize, which this method performs, and until the next run(): they are views of one pinned pool
        this object keeps)."""
        batches = list(batches)
        for bt in batches:
            if len(bt) not in (2, 4):
                raise ValueError('HostPipeline batches are (wav, video) or (wav, video, wav_lengths, video_lengths)')
            if len(bt) == 4 and self.ex.fusion != 'video' and hasattr(self.ex.audio, 'min_frames'):
                nmin = int(bt[2].min())
                if ops.num_frames(nmin, feat_type=self.ex.feat_type) < self.ex.audio.min_frames:
                    raise ValueError('ragged batch holds an utterance of %d samples: shorter than the audio model\'s '
                                     'receptive field (%d feature frames needed)' % (nmin, self.ex.audio.min_frames))
        main = torch.cuda.current_stream(self.device)
        for ev in self._free:
            ev.record(main)
        outs = []
        pool = None
        if batches:
            self._upload(0, *batches[0])
        for k in range(len(batches)):
            slot = k % self.slots
            if k + 1 < len(batches):
                self._upload(k + 1, *batches[k + 1])
            wav_d, vid_d, wl_d, vl_d, ragged = self._stage[slot]
            if not ragged:
                wl_d = vl_d = None
            if self.ex.fusion in ('audio', 'video'):
                main.wait_event(self._ready[slot])
                emb = self.ex.extract(wav_d, vid_d, wl_d, vl_d)
            else:                  # same arithmetic as AVExtractor.extract, split at the two upload events
                main.wait_event(self._ready_wav[slot])
                xv = self.ex.audio_embedding(wav_d, wl_d)
                main.wait_event(self._ready[slot])
                emb = self.ex.fuse(xv, self.ex.video_embedding(vid_d, vl_d))
            self._free[slot].record(main)
            pstream = None
            if post is not None:
                emb = post(emb)
                pstream = getattr(post, 'stream', None)      # dist.OverlappedGather: the result lives on its stream
            if pool is None:       # one pinned pool, kept across runs (cudaHostAlloc per step would serialise the pipeline)
                need = (len(batches),) + tuple(emb.shape)
                pool = self._out_pool
                if pool is None or pool.dtype != emb.dtype or tuple(pool.shape[1:]) != need[1:] or pool.shape[0] < need[0]:
                    pool = torch.empty(need, dtype=emb.dtype, pin_memory=True)
                    self._out_pool = pool
            h = pool[k]
            if pstream is not None:                          # D2H behind the collective, off the compute stream
                with torch.cuda.stream(pstream):
                    h.copy_(emb, non_blocking=True)
            else:
                h.copy_(emb, non_blocking=True)
            outs.append(h)
        torch.cuda.synchronize(self.device)
        return outs


def build_models(device='cuda', audio_arch='etdnn', pooling='statistic', seed=1, randomize=True):
    """Random-init (seeded) audio + video drop-in models in eval mode, as bench/tests use them."""
    from . import synth
    from .audio_models.tdnn import SpeakerEmbNet
    from .video_models.model import Lipreading
    aopts = synth.audio_opts(audio_arch, pooling)
    audio = SpeakerEmbNet(aopts)
    audio.load_state_dict(synth.make_audio_state_dict(aopts, seed=seed, randomize=randomize))
    video = Lipreading(relu_type='prelu', backbone_type='resnet', extract_feats=True, tcn_options=synth.TCN_OPTIONS)
    video.load_state_dict(synth.make_video_state_dict(seed=seed, randomize=randomize))
    return audio.to(device).eval(), video.to(device).eval()
