"""Data-parallel gradient exchange of the training step, overlapped with the backward pass.

The reference's local DDP (megatron/model/distributed.py:35-63, called from training.py:165-199) waits for
the whole backward pass, flattens every gradient into one buffer, all-reduces it, and copies it back.  Here
gradients are BORN in flat buffers: every parameter's `.grad` is a view into its bucket's contiguous buffer
(so there is nothing to flatten or copy back, and one memset per bucket replaces ~600 zero fills), buckets are
laid out in reverse registration order — the order in which backward produces gradients — and a bucket's
NCCL all-reduce is launched asynchronously from an autograd hook the moment its last gradient has been
accumulated, so it runs over NVLink underneath the rest of the backward pass.  `finish()` launches whatever
is left (buckets whose parameters received no gradient this step) and makes the compute stream wait.
With `main_grad=True` the buffers are fp32 sinks the backward kernels add into directly (see the class).

Parameters themselves can be re-homed the same way (`flatten_parameters`), which turns the optimizer's
master-weight copy-back into one copy per bucket.  torch.distributed (NCCL) is the transport; nothing here
touches the kernels.
"""
import torch


class _Bucket(object):
    __slots__ = ("params", "offsets", "grad", "pending", "work", "numel")


class GradientBuckets(object):
    """main_grad=True: the flat buffers are fp32 and are attached to the parameters as `main_grad` sinks that the
    backward kernels of emdr2_b200/autograd.py accumulate into directly (no `.grad` tensors, no zero fills, casts
    or AccumulateGrad adds; the optimizer's fp32 master gradients ARE these buffers).  A gradient that still arrives
    through `.grad` (a parameter used by a plain torch op) is folded into the sink by the hook."""

    def __init__(self, params, group=None, bucket_bytes=64 << 20, main_grad=False):
        import torch.distributed as dist
        self.main_grad = bool(main_grad)
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        params = [p for p in params if p.requires_grad]
        self.params = params
        self.buckets = []
        self._bucket_of = {}
        self._handles = []
        cur, size = [], 0
        for p in reversed(params):                       # backward reaches the last layers first
            nbytes = p.numel() * (4 if self.main_grad else p.element_size())
            if cur and (size + nbytes > bucket_bytes or p.dtype != cur[0].dtype or p.device != cur[0].device):
                self._seal(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self._seal(cur)
        for p in params:
            self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.launched = 0

    def _seal(self, group_params):
        b = _Bucket()
        b.params, b.offsets, off = list(group_params), [], 0
        for p in group_params:
            b.offsets.append(off)
            off += -(-p.numel() // 8) * 8                # 16-byte aligned views
        b.numel = off
        dtype = torch.float32 if self.main_grad else group_params[0].dtype
        b.grad = torch.zeros(off, dtype=dtype, device=group_params[0].device)
        for p, o in zip(b.params, b.offsets):
            view = b.grad[o:o + p.numel()].view_as(p)
            if self.main_grad:
                p.main_grad, p._on_main_grad, p.grad = view, self._on_main_grad, None
            else:
                p.grad = view
            self._bucket_of[p] = b
        b.pending, b.work = len(b.params), None
        self.buckets.append(b)

    # ------------------------------------------------------------------ per step
    def start_step(self):
        """Zero the flat buffers (one memset each) and re-arm the hooks.  Call before the forward pass."""
        self.launched = 0
        for b in self.buckets:
            b.grad.zero_()
            b.pending, b.work = len(b.params), None
            for p, o in zip(b.params, b.offsets):
                if self.main_grad:
                    p._pending_main_grads, p.grad = 0, None
                elif p.grad is None or p.grad.data_ptr() != b.grad.data_ptr() + o * b.grad.element_size():
                    p.grad = b.grad[o:o + p.numel()].view_as(p)      # an optimizer or a caller dropped the view

    def _on_grad(self, p):
        if self.main_grad:                               # arrived through autograd's .grad: fold it into the sink
            if p.grad is not None:
                p.main_grad.add_(p.grad)
                p.grad = None
        self._complete(p)

    def _on_main_grad(self, p):                          # the last kernel contribution of this step has been enqueued
        self._complete(p)

    def _complete(self, p):
        b = self._bucket_of[p]
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b):
        if self.world > 1 and b.work is None:
            import torch.distributed as dist
            b.work = dist.all_reduce(b.grad, group=self.group, async_op=True)
            self.launched += 1

    def finish(self):
        """All-reduce what has not been launched yet, then order the compute stream after every all-reduce.
        Gradients hold the SUM over ranks (divide by `world` where they are consumed)."""
        for b in self.buckets:
            if b.work is None:
                self._launch(b)
        for b in self.buckets:
            if b.work is not None:
                b.work.wait()

    def close(self):
        for h in self._handles:
            h.remove()
        self._handles = []


def flatten_parameters(buckets):
    """Re-home every bucket's parameters as views of one flat buffer (same layout as the gradient buffer).
    Returns the list of flat parameter buffers, index-aligned with `buckets.buckets`."""
    flats = []
    with torch.no_grad():
        for b in buckets.buckets:
            flat = torch.zeros(b.numel, dtype=b.params[0].dtype, device=b.grad.device)
            for p, o in zip(b.params, b.offsets):
                view = flat[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
            flats.append(flat)
    return flats
