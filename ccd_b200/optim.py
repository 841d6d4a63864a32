"""AdamW with the caller-side step glue of train.py fused into ONE pass over the parameter state.

`AdamW` is a drop-in for `torch.optim.AdamW(params_groups)` as train.py:131-133 builds it (same param_groups / state
layout: 'step', 'exp_avg', 'exp_avg_sq'; state_dict round-trips with torch.optim.AdamW).  `step()` alone is the plain
optimizer step; `step(clip_grad=c, ema=TeacherEMA, ema_momentum=m)` additionally folds in
    utils.clip_gradients(student, c)        Dino/modules/utils.py:132-141
    the teacher EMA loop                    train.py:264-272
and refreshes the bf16 GEMM-operand copies of student and teacher (encoder / head `bf16_copies()`), so the sequence
clip -> optimizer.step -> EMA -> (next forward's weight casts) is two kernel launches (squared norms; fused update)
instead of ~10 launches and four extra passes over 46 M parameters.  sm_100a kernels only: CUDA parameters required.

The pointer table of the kernels is built once per set of participating parameters; gradient addresses (which move
between steps: gradients are fresh allocations after zero_grad(set_to_none=True)) go through a small per-step pointer
array uploaded asynchronously from pinned memory, so a step never rebuilds or re-uploads the table.
"""
import torch

from . import ops

CHUNK = 1 << 14


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._plans = {}

    # ---- state_dict interchange with torch.optim.AdamW: 'step' is a tensor there, a python int here ----
    def state_dict(self):
        # torch's Optimizer.state_dict() hands out the LIVE per-parameter dicts: build copies, never write into them
        sd = super().state_dict()
        sd["state"] = {k: ({**st, "step": torch.tensor(float(st["step"]))}
                           if "step" in st and not torch.is_tensor(st["step"]) else dict(st))
                       for k, st in sd["state"].items()}
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for st in self.state.values():
            if torch.is_tensor(st.get("step")):
                st["step"] = int(st["step"].item())
        self._plans.clear()

    def _build_plan(self, todo, ema_pairs, bf16_of, dev, want_norms):
        """todo = [(group_index, step, [params])]: one launch per entry (parameters of a launch share lr / betas / step)."""
        n_grad = sum(len(ps) for _, _, ps in todo)
        if n_grad >= 1 << 16:
            raise RuntimeError("ccd_b200.optim.AdamW: more than 65535 parameter tensors")
        norms = torch.zeros(max(1, n_grad), dtype=torch.float32, device=dev) if want_norms else None
        launches, covered = [], set()
        updated, recast = set(), set()          # ids of tensors rewritten by the pass / whose bf16 copy is rewritten too
        teacher_of = {id(s): t for s, t in ema_pairs}
        idx = 0
        for gi, step, ps in todo:
            rows = []
            wd_flag = 1 if self.param_groups[gi]["weight_decay"] != 0 else 0
            for p in ps:
                st = self.state[p]
                if not (p.is_cuda and p.is_contiguous() and p.dtype == torch.float32):
                    raise RuntimeError("ccd_b200.optim.AdamW needs contiguous fp32 CUDA parameters")
                t = teacher_of.get(id(p))
                pb = bf16_of.get(id(p))
                tb = bf16_of.get(id(t)) if t is not None else None
                sq = norms.data_ptr() + 4 * idx if want_norms else 0
                covered.add(id(p))
                updated.add(id(p))
                if pb is not None:
                    recast.add(id(p))
                if t is not None:
                    updated.add(id(t))
                    if tb is not None:
                        recast.add(id(t))
                n = p.numel()
                for o in range(0, n, CHUNK):
                    rows.append((p.data_ptr() + 4 * o, idx | (o << 16), st["exp_avg"].data_ptr() + 4 * o,
                                 st["exp_avg_sq"].data_ptr() + 4 * o, sq, t.data_ptr() + 4 * o if t is not None else 0,
                                 pb.data_ptr() + 2 * o if pb is not None else 0, tb.data_ptr() + 2 * o if tb is not None else 0,
                                 min(CHUNK, n - o), wd_flag))
                idx += 1
            launches.append([gi, rows])
        # teacher pairs whose student parameter gets no AdamW update this step (no gradient / frozen): EMA only
        extra = []
        for s, t in ema_pairs:
            if id(s) in covered:
                continue
            tb = bf16_of.get(id(t))
            updated.add(id(t))
            if tb is not None:
                recast.add(id(t))
            n = s.numel()
            for o in range(0, n, CHUNK):
                extra.append((s.data_ptr() + 4 * o, 0, 0, 0, 0, t.data_ptr() + 4 * o, 0, tb.data_ptr() + 2 * o if tb is not None else 0,
                              min(CHUNK, n - o), 2))
        if extra:
            if launches:
                launches[0][1] = launches[0][1] + extra
            else:
                launches.append([0, extra])
        out, keep = [], []
        for gi, rows in launches:
            t, h = ops._upload_table(rows, dev)
            out.append((gi, t, len(rows)))
            keep.append(h)
        return {"launches": out, "norms": norms, "keep": keep, "updated": updated, "recast": recast,
                "grad_ptrs": torch.zeros(max(1, n_grad), dtype=torch.int64, device=dev), "last_ptrs": None}

    @torch.no_grad()
    def step(self, closure=None, clip_grad=0.0, ema=None, ema_momentum=0.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        ema_pairs = ema.pairs if ema is not None else []
        mods = ema.bf16_modules if ema is not None else []
        bf16_of = {}
        for m in mods:
            for p, b in m.bf16_copies():
                bf16_of[id(p)] = b
        todo, key, gptrs, dev = [], [], [], None
        for gi, group in enumerate(self.param_groups):
            by_step = {}
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if g.is_sparse or not (g.is_cuda and g.is_contiguous() and g.dtype == torch.float32):
                    raise RuntimeError("ccd_b200.optim.AdamW needs dense contiguous fp32 CUDA gradients")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                by_step.setdefault(st["step"], []).append(p)
                dev = p.device
            # a launch = the parameters of one group that share a step count (bias corrections are launch scalars); the
            # table depends on that PARTITION, not on the counts themselves
            for rank, (step, ps) in enumerate(sorted(by_step.items())):
                todo.append((gi, step, ps))
                key.append((gi, rank, tuple(id(p) for p in ps)))
                gptrs += [p.grad.data_ptr() for p in ps]
        if dev is None:
            if not ema_pairs:
                return loss
            dev = ema_pairs[0][0].device
        want_norms = clip_grad is not None and clip_grad > 0
        key = (tuple(key), tuple(b.data_ptr() for b in bf16_of.values()), want_norms, len(ema_pairs))
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= 8:
                self._plans.clear()
            plan = self._plans[key] = self._build_plan(todo, ema_pairs, bf16_of, dev, want_norms)
        if gptrs and gptrs != plan["last_ptrs"]:
            # pinned staging + async copy: no host stall; the caching host allocator keeps the block alive until the copy ran
            plan["grad_ptrs"].copy_(torch.tensor(gptrs, dtype=torch.int64).pin_memory(), non_blocking=True)
            plan["last_ptrs"] = gptrs
        if want_norms and todo:
            plan["norms"].zero_()
            for gi, table, n in plan["launches"]:
                ops.grad_sqnorm(table, plan["grad_ptrs"], n)
        steps = [step for _, step, _ in todo] or [1]
        for (gi, table, n), step in zip(plan["launches"], steps):
            g = self.param_groups[gi]
            b1, b2 = g["betas"]
            ops.fused_adamw(table, plan["grad_ptrs"], n, g["lr"], b1, b2, g["eps"], g["weight_decay"], 1.0 - b1 ** step,
                            1.0 - b2 ** step, float(clip_grad) if want_norms else 0.0, float(ema_momentum))
        touched = [p for _, _, ps in todo for p in ps] + [t for _, t in ema_pairs]
        # this pass vouches for a module's bf16 copies only if it rewrote EVERY one of them (parameter and copy in the same
        # kernel); anything else is re-cast by the module's next forward (encoder.py: freshness rule)
        fresh = [m for m in mods if m.bf16_copies() and all(id(p) in plan["recast"] for p, _ in m.bf16_copies())]
        if touched:
            torch.autograd.graph.increment_version(touched)       # raw-pointer in-place updates
        for m in fresh:
            m.bf16_mark_fresh()
        return loss
