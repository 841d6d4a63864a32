"""Student / teacher wrapper (drop-in for ABIDINOModel, Dino/model/dino_vision.py:21-115).

Same constructor, attributes (.backbone .segmentation .head), forward signature and output dict keys as the
reference, but:
  * the connected-component labelling runs on the GPU (one CTA per image) on a side stream, overlapping the encoder;
    the only host round trip left is the 4-byte read of the ragged row count R that the [2R,K] output shape needs
    (the reference moves masks to the host, labels them in a Python loop and uploads a float64 [B,26,32,128] array);
  * cluster maps are one uint32 bitmask per pixel for both views (ClusterMaps); the dense [2B,26,32,128] fp32 tensor
    of the reference is materialised only if somebody asks for it (ClusterMaps.dense()), e.g. a reference teacher;
  * pooling writes only the selected rows.
"""
import torch
import torch.nn as nn

from . import ops

SLOTS = 26


class _RowCount:
    """The single device->host read of the step: R = offs[-1], the ragged row count that fixes the [2R,K] output shape.
    The 4-byte copy into pinned memory is queued right behind the plan kernels; the host only waits for it when the
    value is first needed (CharPoolFn.forward), i.e. after the whole encoder forward has been enqueued, so neither the
    host nor the GPU idles on the round trip."""
    _ring, _next = [], 0

    def __init__(self, offs):
        cls = _RowCount
        if len(cls._ring) < 8:
            cls._ring.append(torch.empty(1, dtype=torch.int32).pin_memory())
        self.host = cls._ring[cls._next % len(cls._ring)]
        cls._next += 1
        self.host.copy_(offs[-1:], non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()
        self.value = None

    def get(self):
        if self.value is None:
            self.event.synchronize()
            self.value = int(self.host[0])
        return self.value


class ClusterMaps:
    """What `student_output['zero']` holds: per-pixel slot bitmasks of both views + the pooling plan.
    train.py:233 passes it straight to the teacher (`clusters=student_output['zero']`)."""

    def __init__(self, bits, tot4, cnt, offs, new_index_u8, R, ready_event=None):
        self.bits, self.tot4, self.cnt, self.offs, self.new_index_u8, self._R = bits, tot4, cnt, offs, new_index_u8, R
        self.ready_event = ready_event
        self._dense = None

    @property
    def R(self):
        return self._R.get() if isinstance(self._R, _RowCount) else self._R

    @property
    def shape(self):
        return torch.Size((self.bits.shape[0], SLOTS, 32, 128))

    def dense(self):
        """The reference's float tensor [2B,26,32,128] (dino_vision.py:71-78)."""
        if self._dense is None:
            self._dense = ops.bits_to_dense(self.bits)
        return self._dense

    @staticmethod
    def from_dense(dense):
        bits = ops.dense_to_bits(dense.contiguous().float())
        tot4, cnt, offs, new_index = ops.char_plan(bits)
        return ClusterMaps(bits, tot4, cnt, offs, new_index, _RowCount(offs))


class CharPoolFn(torch.autograd.Function):
    """Mask-guided character pooling + ragged select (dino_vision.py:38-49, 80-87) -> rows [2R,E]."""

    @staticmethod
    def forward(ctx, tokens, cm):
        n2, _, E = tokens.shape
        rows = ops.char_pool_fwd(tokens.contiguous().view(n2 * 256, E), cm.bits, cm.tot4, cm.cnt, cm.offs, cm.R)
        ctx.cm, ctx.E, ctx.n2 = cm, E, n2
        return rows

    @staticmethod
    def backward(ctx, d_rows):
        cm = ctx.cm
        d_tok = ops.char_pool_bwd(d_rows.contiguous().float(), cm.bits, cm.tot4, cm.cnt, cm.offs, ctx.E)
        return d_tok.view(ctx.n2, 256, ctx.E), None


class ABIDINOModel(nn.Module):
    def __init__(self, backbone, Segmentation, head):
        super().__init__()
        backbone.fc, backbone.head = nn.Identity(), nn.Identity()       # dino_vision.py:33
        self.backbone = backbone
        self.segmentation = Segmentation
        self.head = head
        self._side = None

    def _segments(self, x, metrics, target_mask, epoch, seg_logits_fn):
        """CCL + warp + pooling plan on a side stream (independent of the encoder while epoch < 30)."""
        cur = torch.cuda.current_stream()
        if epoch < 30:                                                   # dino_vision.py:59-63 dataset mask
            if self._side is None:
                self._side = torch.cuda.Stream()
            side = self._side
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                cm = self._plan(target_mask.contiguous().float(), 0, metrics, target_mask.shape[0])
                ev = torch.cuda.Event()
                ev.record(side)
            cm.ready_event = ev
            for t in (cm.bits, cm.tot4, cm.cnt, cm.offs, cm.new_index_u8):
                t.record_stream(cur)
            target_mask.record_stream(side)
            metrics.record_stream(side)
            return cm
        seg = seg_logits_fn()                                            # :64-70 self-predicted mask of view 1
        return self._plan(seg.detach().contiguous().float(), 1, metrics, seg.shape[0] // 2)

    @staticmethod
    def _plan(src, mode, metrics, n_view):
        bits1, _, _ = ops.ccl_label(src, mode, n_view)
        bits2 = ops.warp_bits(bits1, metrics.contiguous().float())
        bits = torch.cat([bits1, bits2])
        tot4, cnt, offs, new_index = ops.char_plan(bits)
        return ClusterMaps(bits, tot4, cnt, offs, new_index, _RowCount(offs))     # R is read lazily (see _RowCount)

    def forward(self, x, metrics, target_mask, epoch, clusters=None, index=None):
        if not x.is_cuda:
            raise RuntimeError("ccd_b200.ABIDINOModel runs on CUDA (sm_100a) only; there is no CPU path")
        views = torch.cat([x[:, 1], x[:, 2]])                            # dino_vision.py:52-54 (view 0 is never used)
        student = clusters is None
        cm = None
        if student and epoch < 30:
            cm = self._segments(x, metrics, target_mask, epoch, None)    # launched before the encoder is enqueued
        tokens, taps = self.backbone(views)
        n2, _, E = tokens.shape
        if student:
            seg_out = self.segmentation(taps)
            if cm is None:
                cm = self._segments(x, metrics, target_mask, epoch, lambda: seg_out)
            elif cm.ready_event is not None:
                torch.cuda.current_stream().wait_event(cm.ready_event)
            rows = CharPoolFn.apply(tokens, cm)
            logits = self.head(rows)
            return {"instances_view": logits, "mask": seg_out, "image": x, "zero": cm, "index": cm.new_index_u8.bool()}
        cm = clusters if isinstance(clusters, ClusterMaps) else ClusterMaps.from_dense(clusters)
        if cm.ready_event is not None:
            torch.cuda.current_stream().wait_event(cm.ready_event)
        rows = CharPoolFn.apply(tokens, cm)
        logits = self.head(rows)
        return {"instances_view": logits, "feature": tokens.view(n2, 8, 32, E).permute(0, 3, 1, 2)}
