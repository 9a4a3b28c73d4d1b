"""Fused execution plan for DYNAMIC quantisation (reference QuantAct.forward with dynamic=True,
nemo/quantization/utils/quant_modules.py:149-167; BASELINE config 4: W4/A8, ranges from the batch itself).

In dynamic mode every QuantAct's scale is a function of the tensor it receives, so nothing can be folded on the
host.  This plan keeps the whole forward on the device and int8-resident anyway:

  per conv -> consumer QuantAct                                            kernels
    statistics pass     conv with QASR_EPI_STATS (no tensor written)       dwconv / tcgen05 GEMM
    exact absmax        rigorous bounds on max |y| per channel pick the    qasr_dyn_refine
                        candidate channels; those again, position by
                        position, with the reference's float noise
                        (quant_modules.py:301)
    range -> scale ->   (m, e) of the producer, s_out / b_int of the       qasr_dyn_finalize
                        consumer convs; under data parallelism the ranks'
                        absmax values meet here through peer memory
    requantising pass   the same conv with the literal epilogue            dwconv / tcgen05 GEMM
  res_act (jasper.py:681): both convs of the block tail in ONE kernel for the statistics and the requantising
  pass; its int8 output goes through an integer min / max (qasr_minmax_i8) and a LUT requantisation with device
  side (m, e) (qasr_requant_i8_dev) into the next block's input codes.

All of it is recorded into one CUDA graph.  Topologies: a residual branch is one 1x1 conv from the block input
(QuartzNet); Jasper's dense-residual chains run on the module-level path.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops


class _Codes:
    """int8 activation codes resident in HBM: [B][C][ld], their QuantAct's scale lives in `scale` (device scalar)."""

    def __init__(self, tensor, C_, T, unsigned, scale, stage):
        self.tensor, self.C, self.T, self.unsigned, self.scale, self.stage = tensor, C_, T, unsigned, scale, stage


class _Conv:
    """Static integer parameters of one QuantConv1d + its device-filled epilogue table."""

    def __init__(self, conv, device):
        p = conv._prepare()
        self.module = conv
        self.depthwise = p["depthwise"]
        self.C_out, self.C_in = conv.out_channels, conv.in_channels
        self.k, self.stride = conv.kernel_size[0], conv.stride[0]
        self.dil, self.pad = conv.dilation[0], conv.padding[0]
        if not self.depthwise and (self.k > 1 or self.stride > 1):
            raise NotImplementedError("dynamic plan: dense k > 1 / strided convolutions (Jasper) run module by module")
        w_int = p["w_int"].detach()
        self.w_raw = w_int.to(torch.int8).contiguous()
        self.wabs = w_int.abs().reshape(self.C_out, -1).sum(dim=1).to(torch.int32).contiguous()
        self.s_w = p["s_w"].detach().float().contiguous()
        self.bias = None if p["bias"] is None else p["bias"].detach().float().contiguous()
        self.w8, self.packed = p["w8"], p["packed"]
        self.table = torch.zeros((self.C_out, 4), dtype=torch.int32, device=device)       # qasr_chan records
        self.stats = torch.zeros((self.C_out, 2), dtype=torch.int64, device=device)       # qasr_stats records
        self.cand = torch.full((8,), -1, dtype=torch.int32, device=device)


class DynamicPlan:
    def __init__(self, encoder, decoder, in_shape, device, use_graph=True, dp=None):
        """dp: None, or dict(rank=, world=, slots=<PeerSlots>) for a batch sharded over the ranks of one node."""
        self.device = torch.device(device)
        self.B, self.C_in, self.T = (int(v) for v in in_shape)
        self.encoder, self.decoder = encoder, decoder
        self.dp = dp
        self.ops, self._keep = [], []
        self.stats = dict(kernels=0, convs=0, finalize=0)
        self.x_in = torch.zeros((self.B, self.C_in, self.T), dtype=torch.float32, device=self.device)
        self.lens = [torch.zeros(self.B, dtype=torch.int32, device=self.device)]
        self._lens_fns, self._len_stages = [], []
        self._convs = []
        self._n_slots = 0
        self.finalize_args = {}      # op index -> qasr_dyn_finalize_args (its `phase` is read at launch time)
        self.cur_T = self.T
        self._build()
        self.cur_T_out = self.out_frames(self.T)
        self.graph, self.use_graph = None, use_graph

    # -- helpers ---------------------------------------------------------------------------
    def _scalar(self, dtype=torch.float32, n=1, fill=0):
        t = torch.full((n,), fill, dtype=dtype, device=self.device)
        self._keep.append(t)
        return t

    def _add(self, name, fn, keep=None):
        self.ops.append((name, fn))
        self._keep.append(keep)
        self.stats["kernels"] += 1

    def _conv(self, module):
        c = _Conv(module, self.device)
        self._convs.append(c)
        self.stats["convs"] += 1
        return c

    def _new_len_stage(self, stage, k, s, d, p):
        prev = self.lens[stage]
        self._len_stages.append((k, s, d, p))
        self.lens.append(torch.zeros_like(prev))
        out = self.lens[-1]

        def fn(prev=prev, out=out):
            out.copy_(torch.div(prev + (2 * p - d * (k - 1) - 1), s, rounding_mode="floor") + 1)
        self._lens_fns.append(fn)
        return len(self.lens) - 1

    def out_frames(self, T):
        for (k, s, d, p) in self._len_stages:
            T = (T + 2 * p - d * (k - 1) - 1) // s + 1
        return T

    @staticmethod
    def _act_cfg(acts, masked):
        bits = {int(a.activation_bit) for a in acts}
        if len(bits) != 1:
            raise NotImplementedError("consumers of one tensor with different activation bit widths")
        for a in acts:
            if a.per_channel or a.running_stat or not a.dynamic or a.percentile:
                raise NotImplementedError("dynamic plan needs fixed, dynamic, min/max, per-tensor QuantAct modules")
        return bits.pop(), masked

    # -- symbolic walk (conv_asr.py:194-206, jasper.py:641-692) ------------------------------
    def _build(self):
        from nemo.collections.asr.parts.jasper import MaskedConv1d
        blocks = list(self.encoder.encoder_layers)
        if self.decoder is None:
            raise NotImplementedError("dynamic plan is compiled for encoder + decoder")

        def first_conv(block):
            return next(l for l in block.mconv if isinstance(l, MaskedConv1d))

        def consumers_after(bi):
            """(QuantAct list, conv module list, masked) of whatever reads block bi's output."""
            if bi + 1 < len(blocks):
                nb = blocks[bi + 1]
                mods = [first_conv(nb)]
                if nb.res is not None:
                    if nb.dense_residual or len(nb.res) != 1 or len(nb.res[0]) != 1:
                        raise NotImplementedError("dense residual chains run on the module-level path")
                    mods.append(nb.res[0][0])
                return [m.act for m in mods], [m.conv for m in mods], True
            return [self.decoder.act], [self.decoder.decoder_layers[0]], False

        # ---- first layer: range of the float input (quant_modules.py:150-156), SymmetricQuantFunction (:180-184)
        mc0 = first_conv(blocks[0])
        bit0, _ = self._act_cfg([mc0.act], True)
        if bit0 > 8:
            raise NotImplementedError("first-layer activation_bit > 8")
        conv_objs = {}

        def conv_of(module):
            if id(module) not in conv_objs:
                conv_objs[id(module)] = self._conv(module)
            return conv_objs[id(module)]

        mm = self._scalar(n=2)
        scale0 = self._scalar()
        ws = torch.empty(max(int(L.lib().qasr_minmax_workspace_bytes()), 16), dtype=torch.uint8, device=self.device)
        lib, x_in, B, Cin, T = L.lib(), self.x_in, self.B, self.C_in, self.T
        lens0 = self.lens[0]
        self._add("minmax_input", lambda: L.check(lib.qasr_minmax_f32(L.ptr(x_in), None, T, B, Cin, T, L.ptr(lens0), L.ptr(mm),
                                                                      L.ptr(ws), L.stream_ptr())), keep=ws)
        self._finalize(1, mm, None, bit0, scale0, prods=[], prev=None, cons=[conv_of(mc0.conv)])
        codes0 = ops.alloc_codes(B, Cin, T, self.device)
        self._add("quantize_input", lambda: L.check(lib.qasr_quantize_f32(L.ptr(x_in), T, L.ptr(scale0), bit0, L.ptr(codes0),
                                                                          codes0.shape[-1], None, B, Cin, T, L.ptr(lens0),
                                                                          L.stream_ptr())))
        cur = _Codes(codes0, Cin, T, False, scale0, 0)

        for bi, block in enumerate(blocks):
            block_in = cur
            layers = list(block.mconv)
            mcs = [(i, l) for i, l in enumerate(layers) if isinstance(l, MaskedConv1d)]
            for l in layers:
                if not isinstance(l, (MaskedConv1d, nn.ReLU, nn.Dropout)):
                    raise NotImplementedError("layer %r in a dynamic plan (call bn_folding() first)" % (l,))
            if not isinstance(list(block.mout)[0], nn.ReLU):
                raise NotImplementedError("block activation %r" % (list(block.mout)[0],))
            has_res = block.res is not None
            if has_res and (block.dense_residual or len(block.res) != 1 or len(block.res[0]) != 1):
                raise NotImplementedError("dense residual chains run on the module-level path")
            for j, (li, mc) in enumerate(mcs):
                last = j == len(mcs) - 1
                conv = conv_of(mc.conv)
                T_new = (cur.T + 2 * conv.pad - conv.dil * (conv.k - 1) - 1) // conv.stride + 1
                stage = cur.stage
                if conv.stride != 1 or T_new != cur.T:
                    stage = self._new_len_stage(cur.stage, conv.k, conv.stride, conv.dil, conv.pad)
                if last and has_res:
                    rc = block.res[0][0]
                    if rc.conv.stride[0] != 1 or rc.conv.kernel_size[0] != 1 or block_in.T != T_new:
                        raise NotImplementedError("strided residual branch")
                    if int(rc.act.activation_bit) != int(first_conv(block).act.activation_bit):
                        raise NotImplementedError("residual branch quantised with another bit width than the block input")
                    self._act_cfg([rc.act], True)
                    self._act_cfg([block.res_act], False)
                    ra_codes = self._emit_dual(conv, cur, conv_of(rc.conv), block_in, int(block.res_act.activation_bit),
                                               T_new, stage)
                    acts, cmods, masked = consumers_after(bi)
                    bit, _ = self._act_cfg(acts, masked)
                    cur = self._emit_block_output(ra_codes, bit, masked, [conv_of(m) for m in cmods], stage)
                else:
                    relu = (li + 1 < len(layers) and isinstance(layers[li + 1], nn.ReLU)) or last
                    if last:
                        acts, cmods, masked = consumers_after(bi)
                    else:
                        nxt = mcs[j + 1][1]
                        acts, cmods, masked = [nxt.act], [nxt.conv], True
                    bit, _ = self._act_cfg(acts, masked)
                    cur = self._emit_single(conv, cur, relu, bit, masked, [conv_of(m) for m in cmods], T_new, stage)
        # ---- decoder: 1x1 conv with bias on the decoder QuantAct's codes -> float logits (conv_asr.py:270-275)
        dconv = conv_of(self.decoder.decoder_layers[0])
        self.enc_stage = cur.stage
        self.decoder_act_scale = cur.scale      # device scalar: the decoder QuantAct's scale of the last run
        self.logits = torch.zeros((self.B, dconv.C_out, ops.round_up(cur.T, 16)), dtype=torch.float32, device=self.device)
        flags = L.F_DEVICE_TABLES | L.F_BIAS64 | (0 if cur.unsigned else L.F_B_SIGNED)
        ep = ops.make_epilogue(L.EPI_DEQUANT_F32, dconv.table, flags=flags)
        self._launch_conv("decoder", dconv, cur, None, None, self.logits, ep, cur.T)
        self.logits_conv = dconv

    # -- emitters --------------------------------------------------------------------------
    def _dyn_operand(self, conv, src):
        op = L.DynOperand()
        op.x, op.w, op.s_x = src.tensor.data_ptr(), conv.w_raw.data_ptr(), src.scale.data_ptr()
        op.ch, op.wabs = conv.table.data_ptr(), conv.wabs.data_ptr()
        op.x_ld, op.C_in, op.T_in, op.k = src.tensor.shape[-1], conv.C_in, src.T, conv.k
        op.stride, op.dilation, op.pad = conv.stride, conv.dil, conv.pad
        op.depthwise, op.in_unsigned = int(conv.depthwise), int(src.unsigned)
        op._keep = (src.tensor, conv.w_raw, src.scale, conv.table, conv.wabs)
        return op

    def _launch_conv(self, name, conv0, src0, conv1, src1, out, ep, T_out):
        lib, B = L.lib(), self.B
        if conv0.depthwise:
            assert conv1 is None
            x = src0.tensor

            def run():
                L.check(lib.qasr_dwconv1d_i8(L.ptr(x), x.shape[-1], L.ptr(conv0.w8), L.ptr(out), ops.round_up(T_out, 16), B,
                                             conv0.C_out, src0.T, T_out, conv0.k, conv0.stride, conv0.dil, conv0.pad,
                                             C.byref(ep), L.stream_ptr()))
            self._add("dw_k%d_%s" % (conv0.k, name), run, keep=(ep, x, out))
        else:
            g0 = ops.gemm_operand(src0.tensor, src0.T, conv0.packed, conv0.C_in, conv0.k, conv0.stride, conv0.dil, conv0.pad)
            g1 = None
            if conv1 is not None:
                g1 = ops.gemm_operand(src1.tensor, src1.T, conv1.packed, conv1.C_in, conv1.k, conv1.stride, conv1.dil,
                                      conv1.pad)
            op1 = C.byref(g1) if g1 is not None else None

            def run():
                L.check(lib.qasr_conv1d_gemm_i8(C.byref(g0), op1, L.ptr(out), ops.round_up(T_out, 16), B, conv0.C_out, T_out,
                                                C.byref(ep), L.stream_ptr()))
            self._add("gemm_%dx%d_%s" % (conv0.C_in, conv0.C_out, name), run, keep=(ep, g0, g1, out))

    def _finalize(self, kind, src, src_scale, bit, act_scale, prods, prev, cons, prod_srcs=(), stats=None):
        """prods: convs whose (m, e) this range completes (prod_srcs: their input codes, for the fast-path proof);
        cons: convs that read the new codes; stats: (conv, dual) whose statistics records go back to the reset state."""
        a = L.DynFinalizeArgs()
        a.src_kind, a.src = kind, src.data_ptr()
        a.src_scale = src_scale.data_ptr() if src_scale is not None else None
        a.n_levels = 2 ** (bit - 1) - 1
        a.act_scale = act_scale.data_ptr()
        a.n_prod = len(prods)
        for i, c in enumerate(prods):
            a.prod[i], a.prod_C[i] = c.table.data_ptr(), c.C_out
            a.prod_wabs[i], a.prod_qabs[i] = c.wabs.data_ptr(), 255 if prod_srcs[i].unsigned else 128
        if stats is not None:
            a.stats, a.stats_C, a.stats_dual = stats[0].stats.data_ptr(), stats[0].C_out, int(stats[1])
        if prev is not None:
            a.prev_scale, a.prev_me = prev[0].data_ptr(), prev[1].data_ptr()
        a.n_cons = len(cons)
        for i, c in enumerate(cons):
            a.cons[i], a.cons_sw[i], a.cons_C[i] = c.table.data_ptr(), c.s_w.data_ptr(), c.C_out
            a.cons_bias[i] = c.bias.data_ptr() if c.bias is not None else None
        if self.dp is not None and self.dp["world"] > 1:
            slots = self.dp["slots"]
            a.peer_slots, a.rank, a.world = slots.table_ptr(), self.dp["rank"], self.dp["world"]
            a.slot = self._n_slots
            a.seq = slots.seq_ptr(self._n_slots)
            self._n_slots += 1
            if self._n_slots > slots.n_slots:
                raise RuntimeError("range exchange: more QuantAct ranges than slots (%d)" % slots.n_slots)
        lib = L.lib()
        self.finalize_args[len(self.ops)] = a
        self._add("finalize", lambda: L.check(lib.qasr_dyn_finalize(C.byref(a), L.stream_ptr())),
                  keep=(a, src, src_scale, act_scale, prods, prev, cons))
        self.stats["finalize"] += 1

    def _range_of(self, name, conv0, src0, conv1, src1, relu, masked, T_out, stage):
        """statistics pass -> candidates -> exact absmax (device uint32 holding the float's bits)."""
        lib = L.lib()
        lens = self.lens[stage] if masked else None
        flags = L.F_DEVICE_TABLES | (L.F_MASK if masked else 0) | (0 if src0.unsigned else L.F_B_SIGNED)
        if conv1 is not None and not src1.unsigned:
            flags |= L.F_B1_SIGNED
        ep = ops.make_epilogue(L.EPI_STATS, conv0.table, flags=flags, ch1=conv1.table if conv1 is not None else None,
                               out_lens=lens)
        if conv1 is None:    # reset state of the qasr_stats records (dyn_select restores it after every forward)
            conv0.stats[:, 0], conv0.stats[:, 1] = 2 ** 63 - 1, -2 ** 63
        else:
            conv0.stats[:, 0], conv0.stats[:, 1] = 0xffffffff, 0
        self._launch_conv(name + "_stats", conv0, src0, conv1, src1, conv0.stats, ep, T_out)
        o0 = self._dyn_operand(conv0, src0)
        o1 = self._dyn_operand(conv1, src1) if conv1 is not None else None
        p1 = C.byref(o1) if o1 is not None else None
        absmax = self._scalar(torch.int32)
        stats, cand, B = conv0.stats, conv0.cand, self.B
        kind = "dual" if conv1 is not None else ("dw" if conv0.depthwise else "pw")
        self._add("refine_" + kind, lambda: L.check(lib.qasr_dyn_refine(L.ptr(stats), C.byref(o0), p1, int(relu), L.ptr(lens), B,
                                                                conv0.C_out, T_out, L.ptr(cand), L.ptr(absmax),
                                                                L.stream_ptr())), keep=(o0, o1, lens))
        return absmax

    def _emit_single(self, conv, src, relu, bit, masked, cons, T_out, stage):
        """conv -> [ReLU] -> consumer QuantAct (bit, masked): statistics, range, requantising pass."""
        n = 2 ** (bit - 1) - 1
        if n > (255 if relu else 127):
            raise NotImplementedError("activation codes wider than 8 bits")
        absmax = self._range_of("conv", conv, src, None, None, relu, masked, T_out, stage)
        scale = self._scalar()
        self._finalize(0, absmax, None, bit, scale, prods=[conv], prev=None, cons=cons, prod_srcs=[src], stats=(conv, False))
        out = ops.alloc_codes(self.B, conv.C_out, T_out, self.device)
        lens = self.lens[stage] if masked else None
        flags = (L.F_DEVICE_TABLES | L.F_EXACT_F32 | L.F_EXACT_DYAD | L.F_BIAS64 | (L.F_RELU0 if relu else 0) |
                 (L.F_MASK if masked else 0) | (0 if src.unsigned else L.F_B_SIGNED))
        ep = ops.make_epilogue(L.EPI_REQUANT_I8, conv.table, flags=flags, qmin=-n - 1, qmax=n, out_lens=lens)
        self._launch_conv("conv->act", conv, src, None, None, out, ep, T_out)
        return _Codes(out, conv.C_out, T_out, relu, scale, stage)

    def _emit_dual(self, conv, src, rconv, rsrc, ra_bit, T_out, stage):
        """main conv + residual 1x1 conv -> res_act (jasper.py:671-682): unmasked signed codes, before the block ReLU."""
        n = 2 ** (ra_bit - 1) - 1
        if n > 127:
            raise NotImplementedError("res_act wider than 8 bits")
        absmax = self._range_of("res_act", conv, src, rconv, rsrc, False, False, T_out, stage)
        scale = self._scalar()
        self._finalize(0, absmax, None, ra_bit, scale, prods=[conv, rconv], prev=None, cons=[], prod_srcs=[src, rsrc],
                       stats=(conv, True))
        out = ops.alloc_codes(self.B, conv.C_out, T_out, self.device)
        flags = (L.F_DEVICE_TABLES | L.F_EXACT_F32 | L.F_EXACT_DYAD | L.F_BIAS64 | (0 if src.unsigned else L.F_B_SIGNED) |
                 (0 if rsrc.unsigned else L.F_B1_SIGNED))
        ep = ops.make_epilogue(L.EPI_REQUANT_I8, conv.table, flags=flags, ch1=rconv.table, qmin=-n - 1, qmax=n)
        self._launch_conv("res_act", conv, src, rconv, rsrc, out, ep, T_out)
        return _Codes(out, conv.C_out, T_out, False, scale, stage)

    def _emit_block_output(self, ra, bit, masked, cons, stage):
        """res_act codes -> block ReLU (jasper.py:687) -> the next QuantAct: its range is max of relu(q) * s_ra over the
        counted frames, an integer reduction; requantisation through a 256-entry table with device-side (m, e)."""
        n = 2 ** (bit - 1) - 1
        if n > 255:
            raise NotImplementedError("activation codes wider than 8 bits")
        lib, B = L.lib(), self.B
        lens = self.lens[stage] if masked else None
        mm = torch.tensor([2 ** 31 - 1, -2 ** 31], dtype=torch.int32, device=self.device)
        me = self._scalar(torch.int32, n=2)
        scale = self._scalar()
        x, ld, Cn, T = ra.tensor, ra.tensor.shape[-1], ra.C, ra.T
        self._add("minmax_i8", lambda: L.check(lib.qasr_minmax_i8(L.ptr(x), ld, 1, B, Cn, T, L.ptr(lens), L.ptr(mm),
                                                                  L.stream_ptr())), keep=(x, lens, mm))
        self._finalize(2, mm, ra.scale, bit, scale, prods=[], prev=(ra.scale, me), cons=cons)
        out = ops.alloc_codes(B, Cn, T, self.device)
        self._add("requant_i8", lambda: L.check(lib.qasr_requant_i8_dev(L.ptr(x), ld, 1, L.ptr(me), -n - 1, n, L.ptr(out),
                                                                        out.shape[-1], B, Cn, T, L.ptr(lens), L.stream_ptr())),
                  keep=(x, out, me))
        return _Codes(out, Cn, T, True, scale, stage)

    # -- execution -------------------------------------------------------------------------
    def _run_ops(self):
        for fn in self._lens_fns:
            fn()
        for _, fn in self.ops:
            fn()

    def _replay(self):
        if not self.use_graph:
            self._run_ops()
            return
        if self.graph is None:
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                self._run_ops()
            torch.cuda.current_stream(self.device).wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run_ops()
            self.graph = g
        self.graph.replay()

    def load_inputs(self, x, lens):
        B, Cn, T = x.shape
        if (B, Cn, T) != (self.B, self.C_in, self.T):
            # a dynamic range depends on every frame of the batch, zero padding included (masked frames count as 0 in
            # the reference too, but frames past the batch's own T do not exist there): no frame buckets here
            raise ValueError("dynamic plan compiled for input %s, got %s" % ((self.B, self.C_in, self.T), tuple(x.shape)))
        self.x_in.copy_(x, non_blocking=True)
        self.lens[0].copy_(lens.to(torch.int32), non_blocking=True)

    def run(self, x=None, lens=None):
        if x is not None:
            self.load_inputs(x, lens)
        self._replay()
        return self.logits[:, :, :self.cur_T_out]

    def overflowed(self):
        """True when some range had more candidate channels than the refinement covers (QASR_DYN_MAX_CAND); the
        caller then repeats the batch on the module-level path.  One device -> host read."""
        counts = torch.stack([c.cand[L.DYN_MAX_CAND] for c in self._convs])
        return bool((counts > L.DYN_MAX_CAND).any().item())

    def greedy(self):
        return ops.ctc_argmax(self.logits, self.cur_T_out)

    def log_probs(self):
        return ops.log_softmax_tv(self.logits, self.cur_T_out)

    def out_lens(self):
        return self.lens[self.enc_stage]

    @property
    def logits_scale(self):
        """s_out of the decoder conv of the LAST run (device-side table) as a (1, V, 1) float tensor."""
        return self.logits_conv.table[:, 0].contiguous().view(torch.float32).view(1, -1, 1)


class PeerSlots:
    """Peer-mapped slot tables of the range exchange: every rank owns uint64[n_slots][world] in device memory that
    the other ranks of the node map through CUDA IPC (qasr_comm_*); `gather` is how the 64-byte handles travel
    (torch.distributed.all_gather_object in production, a list in the single-process tests)."""

    def __init__(self, rank, world, n_slots, device, gather=None, local_tables=None):
        """local_tables: every rank's table as a tensor of THIS process (all ranks' plans live in one process and run
        in turn on one stream with the post / wait phases of qasr_dyn_finalize: the single-GPU test form)."""
        self.rank, self.world, self.n_slots = rank, world, n_slots
        self._opened, self._own = [], None
        if local_tables is not None:
            self._local = local_tables
            self._table = torch.tensor([t.data_ptr() for t in local_tables], dtype=torch.int64, device=device)
            self._seq = torch.zeros(n_slots, dtype=torch.int32, device=device)
            return
        lib = L.lib()
        own = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        L.check(lib.qasr_comm_alloc(8 * n_slots * world, C.byref(own), handle))
        self._own = own
        handles = gather(bytes(handle))
        self._opened = []
        ptrs = []
        for r in range(world):
            if r == rank:
                ptrs.append(own.value)
            else:
                p = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                L.check(lib.qasr_comm_open(buf, C.byref(p)))
                self._opened.append(p)
                ptrs.append(p.value)
        self._table = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self._seq = torch.zeros(n_slots, dtype=torch.int32, device=device)

    def table_ptr(self):
        return self._table.data_ptr()

    def seq_ptr(self, i):
        return self._seq.data_ptr() + 4 * i

    def close(self):
        lib = L.load_library()
        for p in self._opened:
            lib.qasr_comm_close(p)
        self._opened = []
        if self._own is not None:
            lib.qasr_comm_free(self._own)
            self._own = None
