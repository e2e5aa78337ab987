"""CUDA backend: turns tensor-level op requests into pre-bound launches of ``libdinopose_sm100a.so``.

The engine (``engine.py``) describes a forward / backward pass ONCE per (batch, resolution, mode)
plan as a sequence of backend calls on statically allocated tensors; each call here validates the
tensors, freezes the raw pointers / shapes into ctypes arguments and appends ``(fn, args)`` to a
``Program``.  Running a step is then a tight loop of ctypes calls on the current CUDA stream (and is
CUDA-graph capturable: no allocation, no synchronisation).

The method set is the op vocabulary of the C ABI (include/dinopose.h), one method per entry point.
``tests/emulator.py`` implements the same vocabulary in plain torch so the engine's orchestration
and weight-packing logic can be checked on a CPU-only box; it is test infrastructure and is never
imported by the package.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import GemmArgs, WgradArgs

ACT = {"none": 0, "relu": 1, "gelu": 2}
ROWMAP = {"identity": 0, "patch_tokens": 1, "nchw": 2, "shuffle2x2": 3}


def _p(t):
    return None if t is None else t.data_ptr()


def _chk(t, dtype, name, contiguous=True):
    if t is None:
        return
    if not t.is_cuda:
        raise _lib.DinoPoseError(f"{name}: expected a CUDA tensor (no CPU path exists)")
    if t.dtype != dtype:
        raise _lib.DinoPoseError(f"{name}: expected {dtype}, got {t.dtype}")
    if contiguous and not t.is_contiguous():
        raise _lib.DinoPoseError(f"{name}: expected a contiguous tensor")


class Program:
    """A recorded list of launches; ``run()`` replays them on the current stream."""

    def __init__(self):
        self.calls = []       # (fn, args, name)
        self.keep = []        # keep ctypes structs / tensors alive
        self.meta = []        # per call: kernel family, algorithmic FLOPs / bytes (roofline accounting)
        self.lib = _lib.lib()

    def add(self, name, fn, *args, keep=(), kernel=None, flops=0.0, bytes=0.0, launches=1):
        self.calls.append((fn, args, name))
        self.meta.append({"name": name, "kernel": kernel or name, "flops": float(flops), "bytes": float(bytes),
                          "launches": launches, "side": self._side})
        self.keep.extend(keep)

    # ---- fork / join: launches recorded between ``fork(k)`` ... ``side(False)`` go to side stream k (1 or 2) and run
    # concurrently with what follows on the main stream until ``join(k)`` (small-grid chains such as the z-head MLP, the
    # weight re-packing or the hourglass' 4x4 / 8x8 layers overlap the tensor-core kernels; weight-gradient GEMMs
    # overlap the HBM-bound BatchNorm backward).  ``side(k)`` switches the recording stream while a fork is open and
    # ``sync`` adds a one-way dependency without closing it.  Program order stays a valid serial order: ``run_timed``
    # and ``run_family`` simply ignore the streams.
    _side = False

    def _stream_op(self, op, k, src=0):
        self.calls.append((None, (op, int(k), int(src)), "stream"))
        self.meta.append({"name": op, "kernel": "host:stream", "flops": 0.0, "bytes": 0.0})

    def fork(self, k=1):
        self._stream_op("fork", k)
        self._side = int(k)

    def side(self, flag):
        """False / 0: record on the main stream; True / 1 / 2: on that (open) side stream."""
        self._side = int(flag)

    def join(self, k=1):
        self._side = False
        self._stream_op("join", k)

    def sync(self, which, k=1, src=0):
        """Inside fork k: "side_wait" = side stream k waits for everything enqueued so far on the main stream (src = 0) or
        on the open side stream `src` (a launch moved to stream k depends on a producer there); "main_wait" = the main
        stream waits for side stream k (the fork stays open).  No-ops outside a fork."""
        assert which in ("side_wait", "main_wait")
        self._stream_op(which, k, src)

    def add_callable(self, name, fn):
        """Host-side step (e.g. a torch op on static tensors) recorded in order with the launches."""
        self.calls.append((None, fn, name))
        self.meta.append({"name": name, "kernel": "host:" + name, "flops": 0.0, "bytes": 0.0, "side": self._side})

    def add_mark(self, tag):
        """A named point in the program (e.g. "gradients up to offset N are final"); ``run(on_mark=f)`` calls
        ``f(tag)`` there -- the trainer uses it to start bucketed all-reduces while the backward continues."""
        self.calls.append((None, tag, "mark"))
        self.meta.append({"name": "mark", "kernel": "host:mark", "flops": 0.0, "bytes": 0.0})

    def __len__(self):
        """number of kernel launches of this library per run"""
        return sum(m.get("launches", 1) for (fn, _a, _n), m in zip(self.calls, self.meta) if fn is not None)

    def run_family(self, kernel):
        """Launch only the calls of one kernel family (``meta['kernel']``), host steps and marks skipped: timing aid for
        bench.py (the launches run on whatever the buffers hold; results are garbage by design)."""
        stream = torch.cuda.current_stream().cuda_stream
        n = 0
        for (fn, args, name), meta in zip(self.calls, self.meta):
            if fn is None or meta["kernel"] != kernel:
                continue
            rc = fn(*args, stream)
            if rc != 0:
                _lib.check(rc, name)
            n += meta.get("launches", 1)
        return n

    def run(self, on_mark=None):
        main = torch.cuda.current_stream()
        stream = main.cuda_stream
        sides = {}            # open forks: k -> stream
        for (fn, args, name), meta in zip(self.calls, self.meta):
            if fn is None:
                if name == "mark":
                    if on_mark is not None:
                        on_mark(args)
                elif name == "stream":
                    op, k, src = args
                    if op == "fork":
                        pool = self._side_streams.setdefault(main.device, {})
                        if k not in pool:
                            pool[k] = torch.cuda.Stream(device=main.device)
                        sides[k] = pool[k]
                        sides[k].wait_stream(main)
                    elif k in sides:
                        if op == "side_wait":
                            sides[k].wait_stream(sides[src] if (src and src in sides) else main)
                        elif op == "main_wait":
                            main.wait_stream(sides[k])
                        else:   # join
                            main.wait_stream(sides.pop(k))
                else:
                    sk = meta.get("side")
                    if sk and sk in sides:      # host-side torch step recorded inside a fork: same stream as its consumers
                        with torch.cuda.stream(sides[sk]):
                            args()
                    else:
                        args()
                continue
            sk = meta.get("side")
            rc = fn(*args, sides[sk].cuda_stream if (sk and sk in sides) else stream)
            if rc != 0:
                _lib.check(rc, name)
        for st in sides.values():     # a fork without a join: never leave work dangling on a side stream
            main.wait_stream(st)

    _side_streams = {}


    def run_timed(self):
        """Replay with a CUDA event pair around every launch (on the launching stream); returns a list of
        dicts {name, kernel, ms, flops, bytes}.  Used by bench.py for the per-kernel roofline figures."""
        stream = torch.cuda.current_stream()
        evs = []
        # keep the GPU busy while the host enqueues: an eager launch + two event records cost ~10 us of CPU time, more
        # than many of the kernels last, and the idle gap would be counted inside the event interval
        torch.cuda._sleep(int(30_000 * len(self.calls)))
        for (fn, args, name), meta in zip(self.calls, self.meta):
            if fn is None:
                if name not in ("mark", "stream"):
                    args()
                continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rc = fn(*args, stream.cuda_stream)
            e1.record(stream)
            if rc != 0:
                _lib.check(rc, name)
            evs.append((meta, e0, e1))
        torch.cuda.synchronize()
        return [dict(m, ms=e0.elapsed_time(e1)) for m, e0, e1 in evs]


class CudaBackend:
    """Records launches of the sm_100a kernels into a ``Program``."""

    name = "cuda"

    def __init__(self):
        self.lib = _lib.lib()
        self.prog = None

    def begin(self):
        self.prog = Program()
        return self.prog

    # ------------------------------------------------------------------ GEMM family
    def gemm(self, A, W, out, *, M, N, K, lda=None, ldw=None, ldo=None, out_dtype="bf16", bias=None, scale=None,
             ls=None, residual=None, ldr=None, aux_out=None, aux_in=None, ld_aux=0, act="none", row_map="identity",
             n_valid=0, map_a=0, map_b=0, conv=None, OH=0, OW=0, NB=0, block_n=0, stats=None, stats_c=0, cta_pair=0,
             ln=None, lora=None, name="gemm"):
        """out = epilogue(A[M,K] @ W[N,K]^T).  ``conv`` = dict(KH, KW, pad, OH, OW) makes A an NHWC
        [NB,IH,IW,C] activation (any strides with unit channel stride) read as an implicit conv.
        ``ln`` = dict(gamma, beta, out, eps): LayerNorm of the output rows fused into the epilogue (row-owning kernel).
        ``lora`` = dict(A, B, scaling, p_drop, seed, y_out, u_out): rank-8 LoRA adapter on the projection output fused
        into the same kernel's epilogue (out = residual + ls * (y + scaling * dropout(y A B)))."""
        _chk(W, torch.bfloat16, name + ".W", contiguous=False)
        a = GemmArgs()
        a.A, a.W, a.out = _p(A), _p(W), _p(out)
        a.M, a.N, a.K = M, N, K
        a.block_n = block_n
        a.ldw = ldw if ldw is not None else W.stride(0)
        if conv is not None:
            _chk(A, torch.bfloat16, name + ".A", contiguous=False)
            assert A.dim() == 4 and A.stride(3) == 1
            a.a_mode = 1
            a.NB, a.IH, a.IW, a.C = A.shape
            a.a_stride_b, a.a_stride_h, a.a_stride_w = A.stride(0), A.stride(1), A.stride(2)
            a.KH, a.KW = conv["KH"], conv["KW"]
            a.pad_y = a.pad_x = conv["pad"]
            a.OH, a.OW = conv["OH"], conv["OW"]
            a.lda = 0
        else:
            _chk(A, torch.bfloat16, name + ".A", contiguous=False)
            a.a_mode = 0
            a.lda = lda if lda is not None else A.stride(0)
            a.OH, a.OW, a.NB = OH, OW, NB
        a.ldo = ldo if ldo is not None else (out.stride(0) if out.dim() == 2 else 0)
        if out.dtype == torch.float32:
            out_dtype = "f32"        # the output tensor decides (pre-BN conv outputs are fp32 in training)
        a.out_dtype = 0 if out_dtype == "bf16" else 1
        _chk(out, torch.bfloat16 if out_dtype == "bf16" else torch.float32, name + ".out", contiguous=False)
        for t, nm in ((bias, "bias"), (scale, "scale"), (ls, "ls")):
            _chk(t, torch.float32, f"{name}.{nm}")
        a.bias, a.scale, a.ls = _p(bias), _p(scale), _p(ls)
        if residual is not None:
            a.residual = _p(residual)
            a.res_is_bf16 = 1 if residual.dtype == torch.bfloat16 else 0
            a.ldr = ldr if ldr is not None else residual.stride(0)
        a.aux_out, a.aux_in, a.ld_aux = _p(aux_out), _p(aux_in), ld_aux
        a.act = ACT[act]
        a.row_map = ROWMAP[row_map]
        a.n_valid, a.map_a, a.map_b = n_valid, map_a, map_b
        _chk(stats, torch.float64, name + ".stats")
        a.stats, a.stats_c = _p(stats), stats_c
        a.cta_pair = cta_pair
        lnk = ()
        if ln is not None:
            _chk(ln["gamma"], torch.float32, name + ".ln_gamma")
            _chk(ln["beta"], torch.float32, name + ".ln_beta")
            _chk(ln["out"], torch.bfloat16, name + ".ln_out", contiguous=False)
            a.ln_gamma, a.ln_beta, a.ln_out = _p(ln["gamma"]), _p(ln["beta"]), _p(ln["out"])
            a.ld_ln, a.ln_eps = ln["out"].stride(0), float(ln.get("eps", 1e-6))
            lnk = (ln["gamma"], ln["beta"], ln["out"])
        if lora is not None:
            for nm in ("A", "B"):
                _chk(lora[nm], torch.float32, f"{name}.lora_{nm}")
            a.lora_A, a.lora_B, a.lora_rank = _p(lora["A"]), _p(lora["B"]), lora["A"].shape[1]
            a.lora_scaling, a.lora_p_drop = float(lora["scaling"]), float(lora.get("p_drop", 0.0))
            a.lora_seed = _p(lora.get("seed"))
            y_out, u_out = lora.get("y_out"), lora.get("u_out")
            _chk(y_out, torch.float32, name + ".lora_y_out", contiguous=False)
            _chk(u_out, torch.float32, name + ".lora_u_out")
            a.lora_y_out, a.ld_lora_y, a.lora_u_out = _p(y_out), (y_out.stride(0) if y_out is not None else N), _p(u_out)
            lnk = (lora["A"], lora["B"], lora.get("seed"), y_out, u_out)
        self.prog.add(name, self.lib.dp_gemm_bf16, C.byref(a), kernel="gemm_kmajor_tcgen05", flops=2.0 * M * N * K,
                      keep=(a, A, W, out, bias, scale, ls, residual, aux_out, aux_in, stats) + lnk)

    def wgrad(self, A, B, out, *, Mc, Nc, so_m, so_n, so_t=0, so_mo=0, so_no=0, m_inner=0, n_inner=0, conv=None,
              P=0, lda=None, ldb=None, block_n=0, splits=0, workspace=None, name="wgrad"):
        """out[off(m)+off(n)+tap*so_t] += sum_p A[p,m] * B[p(+tap),n] (fp32 atomics; out pre-zeroed)."""
        _chk(A, torch.bfloat16, name + ".A", contiguous=False)
        _chk(B, torch.bfloat16, name + ".B", contiguous=False)
        _chk(out, torch.float32, name + ".out", contiguous=False)
        a = WgradArgs()
        a.A, a.B, a.out = _p(A), _p(B), _p(out)
        a.Mc, a.Nc = Mc, Nc
        a.so_m, a.so_mo, a.so_n, a.so_no, a.so_t = so_m, so_mo, so_n, so_no, so_t
        a.m_inner, a.n_inner = m_inner, n_inner
        a.block_n, a.splits = block_n, splits
        if workspace is not None:
            a.workspace, a.workspace_bytes = _p(workspace), workspace.numel() * workspace.element_size()
        if conv is not None:
            assert A.dim() == 4 and B.dim() == 4 and A.stride(3) == 1 and B.stride(3) == 1
            a.mode = 1
            a.NB, a.OH, a.OW = A.shape[0], A.shape[1], A.shape[2]
            a.IH, a.IW = B.shape[1], B.shape[2]
            a.a_sb, a.a_sh, a.a_sw = A.stride(0), A.stride(1), A.stride(2)
            a.b_sb, a.b_sh, a.b_sw = B.stride(0), B.stride(1), B.stride(2)
            a.KH, a.KW = conv["KH"], conv["KW"]
            a.pad_y = a.pad_x = conv["pad"]
        else:
            a.mode = 0
            a.P = P
            a.lda = lda if lda is not None else A.stride(0)
            a.ldb = ldb if ldb is not None else B.stride(0)
            a.KH = a.KW = 1
        pix = (A.shape[0] * A.shape[1] * A.shape[2]) if conv is not None else P
        taps = conv["KH"] * conv["KW"] if conv is not None else 1
        self.prog.add(name, self.lib.dp_wgrad_bf16, C.byref(a), kernel="gemm_wgrad_tcgen05",
                      flops=2.0 * pix * Mc * Nc * taps, keep=(a, A, B, out, workspace), launches=2 if workspace is not None else 1)

    # ------------------------------------------------------------------ backbone row-wise
    def layernorm_fwd(self, x, gamma, beta, y_bf16, y_f32, *, rows, D, T=0, drop_cls=False, eps=1e-6):
        _chk(x, torch.float32, "ln.x", False)
        self.prog.add("layernorm_fwd", self.lib.dp_layernorm_fwd, _p(x), _p(gamma), _p(beta), _p(y_bf16), _p(y_f32),
                      rows, D, T, int(drop_cls), eps, keep=(x, gamma, beta, y_bf16, y_f32), bytes=rows * D * 6.0)

    def layernorm_bwd(self, dy, x, gamma, add_in, dx, *, rows, D, T=0, drop_cls=False, eps=1e-6, ls=None,
                      dx_scaled=None):
        self.prog.add("layernorm_bwd", self.lib.dp_layernorm_bwd, _p(dy), int(dy.dtype == torch.bfloat16), _p(x),
                      _p(gamma), _p(add_in), _p(dx), _p(ls), _p(dx_scaled), rows, D, T, int(drop_cls), eps,
                      keep=(dy, x, gamma, add_in, dx, ls, dx_scaled))

    def patch_im2col(self, px, out, *, B, H, W, Kp):
        _chk(px, torch.float32, "patch_im2col.pixel_values")
        self.prog.add("patch_im2col", self.lib.dp_patch_im2col, _p(px), _p(out), B, H, W, Kp, keep=(px, out))

    def fill_cls(self, x, cls_row, *, B, T, D):
        self.prog.add("fill_cls", self.lib.dp_fill_cls, _p(x), _p(cls_row), B, T, D, keep=(x, cls_row))

    def lora_fwd(self, y, A, Bm, lambda1, x_in, x_out, u_save, *, rows, D, R, scaling, p_drop, seed):
        self.prog.add("lora_fwd", self.lib.dp_lora_fwd, _p(y), _p(A), _p(Bm), _p(lambda1), _p(x_in), _p(x_out),
                      _p(u_save), rows, D, R, scaling, p_drop, _p(seed), keep=(y, A, Bm, lambda1, x_in, x_out, u_save, seed))

    def lora_bwd(self, g, y, u_saved, Bm, lambda1, dA, dB, gu_ws, *, rows, D, R, scaling, p_drop, seed):
        self.prog.add("lora_bwd", self.lib.dp_lora_bwd, _p(g), _p(y), _p(u_saved), _p(Bm), _p(lambda1), _p(dA), _p(dB),
                      _p(gu_ws), rows, D, R, scaling, p_drop, _p(seed),
                      keep=(g, y, u_saved, Bm, lambda1, dA, dB, gu_ws, seed))

    def attention_fwd(self, qkv, ctx, *, B, T, heads, scale):
        _chk(qkv, torch.bfloat16, "attention.qkv")
        self.prog.add("attention_fwd", self.lib.dp_attention_fwd, _p(qkv), _p(ctx), B, T, heads, scale, keep=(qkv, ctx), flops=4.0 * B * T * T * heads * 64, bytes=B * T * heads * 64 * 8.0)

    # ---- backward of an un-frozen encoder layer (Dinov2PoseModel(unfreeze_last_n_layers=n))
    def attention_bwd(self, qkv, ctx, dctx, dqkv, stats, *, B, T, heads, scale):
        for t, nm in ((qkv, "qkv"), (ctx, "ctx"), (dctx, "dctx"), (dqkv, "dqkv")):
            _chk(t, torch.bfloat16, "attention_bwd." + nm)
        _chk(stats, torch.float32, "attention_bwd.stats")
        if stats.numel() < 2 * B * heads * T:
            raise _lib.DinoPoseError("attention_bwd.stats: needs 2*B*heads*T floats")
        self.prog.add("attention_bwd", self.lib.dp_attention_bwd, _p(qkv), _p(ctx), _p(dctx), _p(dqkv), _p(stats), B, T, heads,
                      scale, keep=(qkv, ctx, dctx, dqkv, stats), flops=16.0 * B * T * T * heads * 64, launches=2)

    def layernorm_bwd_params(self, dy, x, dgamma, dbeta, *, rows, D, eps=1e-6):
        _chk(x, torch.float32, "ln_params.x", False)
        self.prog.add("layernorm_bwd_params", self.lib.dp_layernorm_bwd_params, _p(dy), int(dy.dtype == torch.bfloat16), _p(x),
                      _p(dgamma), _p(dbeta), rows, D, eps, keep=(dy, x, dgamma, dbeta))

    def colsum_prod(self, g, a, out, *, P, C):
        _chk(g, torch.float32, "colsum_prod.g")
        _chk(a, torch.bfloat16, "colsum_prod.a")
        self.prog.add("colsum_prod", self.lib.dp_colsum_prod, _p(g), _p(a), _p(out), P, C, keep=(g, a, out))

    # ---- prediction.3 (1x1 conv, 64 -> K <= 32) on CUDA cores, forward and fused backward
    def pred1x1_fwd(self, a, w, bias, out, *, P, HW, C, K):
        _chk(a, torch.bfloat16, "pred1x1.a")
        _chk(w, torch.float32, "pred1x1.w")
        _chk(out, torch.float32, "pred1x1.out")
        self.prog.add("pred1x1_fwd", self.lib.dp_pred1x1_fwd, _p(a), _p(w), _p(bias), _p(out), P, HW, C, K,
                      keep=(a, w, bias, out), flops=2.0 * P * C * K, bytes=P * (C * 2.0 + K * 4.0))

    def pred1x1_bwd(self, g, a, w, d, dW, db, *, P, HW, C, K):
        _chk(g, torch.float32, "pred1x1_bwd.g")
        _chk(a, torch.bfloat16, "pred1x1_bwd.a")
        _chk(d, torch.bfloat16, "pred1x1_bwd.d")
        self.prog.add("pred1x1_bwd", self.lib.dp_pred1x1_bwd, _p(g), _p(a), _p(w), _p(d), _p(dW), _p(db), P, HW, C, K,
                      keep=(g, a, w, d, dW, db), flops=4.0 * P * C * K, bytes=P * (C * 4.0 + K * 4.0))

    def decode(self, hm, idx, xy, conf, *, maps, H, W, target_w, target_h):
        _chk(hm, torch.float32, "decode.heatmaps")
        self.prog.add("decode", self.lib.dp_decode, _p(hm), maps, H, W, float(target_w), float(target_h), _p(idx),
                      _p(xy), _p(conf), keep=(hm, idx, xy, conf), bytes=maps * (H * W * 4.0 + 28.0))

    # ------------------------------------------------------------------ heads
    def im2col(self, x, col, *, NB, IH, IW, C, OH, OW, KH, KW, stride, pad):
        self.prog.add("im2col", self.lib.dp_im2col, _p(x), _p(col), NB, IH, IW, C, OH, OW, KH, KW, stride, pad,
                      keep=(x, col))

    def col2im(self, col, bias, big, *, NB, SH, SW, C, BH, BW, KH, KW, stride, pad):
        self.prog.add("col2im", self.lib.dp_col2im, _p(col), _p(bias), _p(big), int(big.dtype == torch.float32), NB, SH,
                      SW, C, BH, BW, KH, KW, stride, pad, keep=(col, bias, big))

    def dwconv3x3(self, x, w, bias, add, out, *, NB, H, W, C, flip=False):
        self.prog.add("dwconv3x3", self.lib.dp_dwconv3x3, _p(x), _p(w), _p(bias), _p(add), _p(out),
                      int(out.dtype == torch.float32), NB, H, W, C, int(flip), keep=(x, w, bias, add, out))

    def dwconv3x3_wgrad(self, x, dout, dw, *, NB, H, W, C):
        self.prog.add("dwconv3x3_wgrad", self.lib.dp_dwconv3x3_wgrad, _p(x), _p(dout), _p(dw), NB, H, W, C,
                      keep=(x, dout, dw))

    def bn_stats(self, raw, sums, *, P, C):
        self.prog.add("bn_stats", self.lib.dp_bn_stats, _p(raw), int(raw.dtype == torch.float32), _p(sums), P, C,
                      keep=(raw, sums))

    def bn_finalize(self, sums, gamma, beta, rm, rv, scale, shift, mean, invstd, *, C, count, eps=1e-5, momentum=0.1):
        self.prog.add("bn_finalize", self.lib.dp_bn_finalize, _p(sums), _p(gamma), _p(beta), _p(rm), _p(rv), _p(scale),
                      _p(shift), _p(mean), _p(invstd), C, float(count), eps, momentum,
                      keep=(sums, gamma, beta, rm, rv, scale, shift, mean, invstd))

    def bn_fold_eval(self, gamma, beta, rm, rv, conv_bias, scale, shift, *, C, eps=1e-5, mean=None, invstd=None):
        self.prog.add("bn_fold_eval", self.lib.dp_bn_fold_eval, _p(gamma), _p(beta), _p(rm), _p(rv), _p(conv_bias),
                      _p(scale), _p(shift), _p(mean), _p(invstd), C, eps,
                      keep=(gamma, beta, rm, rv, conv_bias, scale, shift, mean, invstd))

    def bn_apply(self, raw, scale, shift, add1, add2, out, *, P, C, relu=True, mode=0):
        self.prog.add("bn_apply", self.lib.dp_bn_apply, _p(raw), int(raw.dtype == torch.float32), _p(scale), _p(shift),
                      _p(add1), _p(add2), _p(out), P, C, int(relu), mode, keep=(raw, scale, shift, add1, add2, out))

    def bn_finalize_apply(self, raw, sums, gamma, beta, rm, rv, scale, shift, mean, invstd, add1, add2, out, *, P, C,
                          relu=True, mode=0, eps=1e-5, momentum=0.1):
        """bn_finalize + bn_apply in one launch (count = P); `sums` is the [9][2*C] fp64 buffer of the layer."""
        self.prog.add("bn_apply", self.lib.dp_bn_finalize_apply, _p(raw), int(raw.dtype == torch.float32), _p(sums), _p(gamma),
                      _p(beta), _p(rm), _p(rv), _p(scale), _p(shift), _p(mean), _p(invstd), _p(add1), _p(add2), _p(out), P, C,
                      int(relu), mode, eps, momentum,
                      keep=(raw, sums, gamma, beta, rm, rv, scale, shift, mean, invstd, add1, add2, out))

    def bn_bwd_reduce(self, dout, raw, add1, scale, shift, mean, invstd, sums, *, P, C, relu=True, mode=0):
        self.prog.add("bn_bwd_reduce", self.lib.dp_bn_bwd_reduce, _p(dout), _p(raw), int(raw.dtype == torch.float32),
                      _p(add1), _p(scale), _p(shift),
                      _p(mean), _p(invstd), _p(sums), P, C, int(relu), mode,
                      keep=(dout, raw, add1, scale, shift, mean, invstd, sums))

    def bn_bwd_apply(self, dout, raw, add1, gamma, scale, shift, mean, invstd, sums, draw, dres, dgamma, dbeta, *, P, C,
                     relu=True, mode=0, eval_mode=False, shuffle_oh=0, shuffle_ow=0):
        self.prog.add("bn_bwd_apply", self.lib.dp_bn_bwd_apply, _p(dout), _p(raw), int(raw.dtype == torch.float32),
                      _p(add1), _p(gamma), _p(scale),
                      _p(shift), _p(mean), _p(invstd), _p(sums), _p(draw), _p(dres), _p(dgamma), _p(dbeta), P, C,
                      int(relu), mode, int(eval_mode), shuffle_oh, shuffle_ow,   # eval_mode 2: frozen statistics
                      keep=(dout, raw, add1, gamma, scale, shift, mean, invstd, sums, draw, dres, dgamma, dbeta),
                      launches=1 if C <= 512 else 2)

    def avgpool2(self, x, out, *, planes, OH, OW):
        self.prog.add("avgpool2", self.lib.dp_avgpool2, _p(x), _p(out), planes, OH, OW, keep=(x, out))

    def hm_grad_to_nhwc(self, g, out, *, NB, K, Kp, OH, OW, up):
        _chk(g, torch.float32, "hm_grad")
        self.prog.add("hm_grad_to_nhwc", self.lib.dp_hm_grad_to_nhwc, _p(g), _p(out), NB, K, Kp, OH, OW, up,
                      keep=(g, out))

    def mean_tokens(self, feat, out, *, B, N, D):
        self.prog.add("mean_tokens", self.lib.dp_mean_tokens, _p(feat), _p(out), B, N, D, keep=(feat, out))

    def mean_tokens_bwd(self, dfeat, dmean, *, B, N, D):
        self.prog.add("mean_tokens_bwd", self.lib.dp_mean_tokens_bwd, _p(dfeat), _p(dmean), B, N, D, keep=(dfeat, dmean))

    def sgemm_small(self, A, sa_m, sa_k, Bm, sb_k, sb_n, Cm, ldc, *, M, N, K, bias=None, relu=False, mask_ref=None,
                    ld_ref=0, p_drop=0.0, seed=None, accumulate=False):
        self.prog.add("sgemm_small", self.lib.dp_sgemm_small, _p(A), sa_m, sa_k, _p(Bm), sb_k, sb_n, _p(Cm), ldc, M, N, K,
                      _p(bias), int(relu), _p(mask_ref), ld_ref, p_drop, _p(seed), int(accumulate),
                      keep=(A, Bm, Cm, bias, mask_ref, seed))

    def relu_mask(self, d, ref, out, *, n, keep_scale=1.0):
        self.prog.add("relu_mask", self.lib.dp_relu_mask, _p(d), _p(ref), _p(out), n, float(keep_scale), keep=(d, ref, out))

    def colsum(self, x, out, *, P, C, ld):
        self.prog.add("colsum", self.lib.dp_colsum, _p(x), int(x.dtype == torch.bfloat16), _p(out), P, C, ld,
                      keep=(x, out))

    # ------------------------------------------------------------------ training step around the model
    def pose_loss(self, hm, thm, kps, z, tz, sums, state, out, scales, dhm, dz, *, B, K, HW, momentum=0.9, rate=0.1):
        for t, nm in ((hm, "heatmaps"), (thm, "target_heatmaps"), (kps, "keypoints"), (z, "z"), (tz, "target_z"),
                      (dhm, "d_heatmaps"), (dz, "d_z")):
            _chk(t, torch.float32, "pose_loss." + nm)
        self.prog.add("pose_loss", self.lib.dp_pose_loss, _p(hm), _p(thm), _p(kps), kps.shape[-1], _p(z), _p(tz), _p(sums),
                      _p(state), _p(out), _p(scales), _p(dhm), _p(dz), B, K, HW, momentum, rate,
                      keep=(hm, thm, kps, z, tz, sums, state, out, scales, dhm, dz), bytes=B * K * HW * 4.0 * 5, launches=3)

    def adamw(self, p, g, m, v, step_dev, *, n, lr, beta1, beta2, eps, weight_decay, grad_scale, hyper=None, bump=True):
        """hyper: optional device fp32 {lr, weight_decay}; when given, lr / weight_decay arguments are ignored and a
        scheduler can change the rate between replays of a captured step."""
        for t, nm in ((p, "params"), (g, "grads"), (m, "exp_avg"), (v, "exp_avg_sq")):
            _chk(t, torch.float32, "adamw." + nm)
        if hyper is not None:
            _chk(hyper, torch.float32, "adamw.hyper")
            self.prog.add("adamw", self.lib.dp_adamw_dev, _p(p), _p(g), _p(m), _p(v), n, _p(hyper), beta1, beta2, eps,
                          grad_scale, _p(step_dev), int(bump), keep=(p, g, m, v, step_dev, hyper), bytes=n * 4.0 * 7,
                          launches=2 if bump else 1)
            return
        assert bump, "slice-wise AdamW needs the device hyper-parameter form (dp_adamw_dev)"
        self.prog.add("adamw", self.lib.dp_adamw, _p(p), _p(g), _p(m), _p(v), n, lr, beta1, beta2, eps, weight_decay,
                      grad_scale, _p(step_dev), keep=(p, g, m, v, step_dev), bytes=n * 4.0 * 7, launches=2)

    def mark(self, tag):
        self.prog.add_mark(tag)

    def fork(self, k=1):
        self.prog.fork(k)

    def side(self, flag):
        self.prog.side(flag)

    def join(self, k=1):
        self.prog.join(k)

    def sync(self, which, k=1, src=0):
        self.prog.sync(which, k, src)

    def pack_weights(self, jobs):
        """jobs: list of (dst bf16 tensor, w fp32 [d0,d1,kh,kw] parameter, order, flips, dst_strides|None): dst (viewed in
        dst order, extents w.shape[order]) = bf16(w with dims `flips` mirrored, permuted by `order`).  One launch."""
        rows, max_total = [], 1
        for dst, w, order, flips, dst_strides in jobs:
            _chk(dst, torch.bfloat16, "pack.dst", contiguous=False)
            _chk(w, torch.float32, "pack.src")
            n = [w.shape[d] for d in order]
            st, off = [], 0
            for d in order:
                sd = w.stride(d)
                if d in flips:
                    off += (w.shape[d] - 1) * sd
                    sd = -sd
                st.append(sd)
            if dst_strides is None:
                dt = [n[1] * n[2] * n[3], n[2] * n[3], n[3], 1]
            else:
                dt = list(dst_strides)
            total = n[0] * n[1] * n[2] * n[3]
            max_total = max(max_total, total)
            rows.append([w.data_ptr(), dst.data_ptr()] + n + st + dt + [off, total])
        table = torch.tensor(rows, dtype=torch.int64).to(jobs[0][0].device)
        self.prog.add("pack_weights", self.lib.dp_pack_weights_bf16, _p(table), len(rows), max_total,
                      keep=(table,) + tuple(j[0] for j in jobs) + tuple(j[1] for j in jobs))

    def add_i64(self, tensors, inc=1):
        """tensors: int64 scalar buffers (BatchNorm num_batches_tracked); *t += inc for all of them in one launch."""
        ptrs = torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64).to(tensors[0].device)
        self.prog.add("add_i64", self.lib.dp_add_i64, _p(ptrs), len(tensors), inc, keep=(ptrs,) + tuple(tensors))

    # ------------------------------------------------------------------ host-side steps on static tensors
    def host(self, name, fn):
        """Record a torch-level step (memset of gradient buffers, seed increment, ...)."""
        self.prog.add_callable(name, fn)
