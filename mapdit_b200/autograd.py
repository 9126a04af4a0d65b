"""Training-mode forward/backward of the whole DiT as one autograd node.

The reference builds its graph from ~4k ATen ops per forward; here the forward saves
exactly the activations the hand-written backward needs and the backward is a fixed kernel schedule
using the closed forms of SURVEY.md §A.3 (weight-norm tangent projection, detached modulate
denominator, mp residual, mp_silu, q/k normalisation, cosine attention).

Gradient flow per block (reverse of src/blocks/dit_block.py:32-37), R = dL/dx residual stream:
  resid_bwd -> fc2 wgrad/dgrad -> mp_silu_bwd -> fc1 wgrad/dgrad -> modulate_bwd (+= R)
  resid_bwd -> out-proj wgrad/dgrad -> attention bwd -> qk_norm_bwd -> qkv wgrad/dgrad -> modulate_bwd (+= R)
"""
import os

import torch

from . import _lib, ops


class _DiTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, t, y, drop_mask, *params):
        out, saved = engine.trainer.forward(x, t, y, drop_mask)
        ctx.engine = engine
        ctx.saved = saved
        return out

    @staticmethod
    def backward(ctx, dout):
        grads = ctx.engine.trainer.backward(ctx.saved, dout.contiguous().float())
        ctx.saved = None
        return (None, None, None, None, None, *grads)


def dit_forward_autograd(engine, x, t, y, drop_mask):
    params = list(engine.m.parameters())
    return _DiTFunction.apply(engine, x, t, y, drop_mask, *params)


class Trainer:
    """Owns the training workspaces (saved activations, gradient scratch) of one Engine."""

    def __init__(self, engine):
        self.e = engine
        self.m = engine.m
        self._bufs = {}
        self.grad_buffers = None  # optional {id(param): preallocated grad tensor} (TrainStep's flat buffer)
        self.grad_hook = None  # optional callable(list_of_(param, grad)) fired as soon as a block's grads are final
        # bf16 path: the weight-gradient GEMMs (+ their weight-norm backward) are issued on a second stream.  They are off the
        # critical path of the backward (nothing downstream reads dW), use 58 registers x 256 threads and ~193 KB of shared
        # memory, so the HBM-bound elementwise backward kernels of the main stream co-reside with them on every SM instead
        # of running alone on an idle tensor pipe.  MAPDIT_WGRAD_STREAM=0 turns it off (A/B in bench.py --wgrad-stream).
        self.wgrad_stream = os.environ.get("MAPDIT_WGRAD_STREAM", "1") != "0"
        # delta = dO.O from the out-proj dgrad epilogue (MAPDIT_EPI_STORE_DELTA) instead of a separate pass over dO and O.  Off by
        # default: measured on DiT-B/2 batch 256 the step does not get faster (43.6 ms with, 43.3 ms without, call c12 of round 2):
        # the epilogue's residual-tile round trips cost the K = 768 dgrad GEMM more than the 35 us kernel it replaces
        self.fuse_delta = os.environ.get("MAPDIT_FUSE_DELTA", "0") == "1"
        self._side = None      # the second stream
        self._side_on = False  # active inside the current backward
        self._side_reads = {}  # scratch buffer name -> event recorded after the side-stream GEMM that last read it
        self._wn_pending = []  # (param, grad buffer holding d(effective weight)) awaiting the multi-tensor weight-norm backward
        self._wn_tables = {}   # pointer signature -> ops.WeightNormBwdBatch (device-resident descriptor table)
        # the saved activations live in ONE workspace per (batch, mode, device): a second train-mode forward overwrites what the
        # first one saved.  Every forward stamps the workspace; a backward whose stamp is stale raises instead of returning
        # gradients of the wrong activations (the reference's autograd keeps both graphs alive).
        self._generation = 0

    # ------------------------------------------------------------------ buffers
    def buffers(self, N, mode, dev):
        key = (N, mode, str(dev))
        B = self._bufs.get(key)
        if B is not None:
            return B
        m = self.m
        D, L, H = m.hidden_size, m.depth, m.num_heads
        T = (m.input_size // m.patch_size) ** 2
        M = N * T
        Hm = m.blocks[0].mlp.hidden_dim
        adt = torch.bfloat16 if mode == "bf16" else torch.float32
        a = dict(device=dev, dtype=adt)
        f = dict(device=dev, dtype=torch.float32)
        ppc2 = m.final_layer.linear.weight.shape[0]
        K1 = m.x_embedder.weight.shape[1]
        W = L * self.e.layout["width"] + 2 * D
        B = dict(
            e=torch.empty(N, 256, **f), t1=torch.empty(N, D, **f), t1s=torch.empty(N, D, **f), temb=torch.empty(N, D, **f),
            yemb=torch.empty(N, D, **f), c=torch.empty(N, D, **f), cs=torch.empty(N, D, **f),
            cs16=torch.empty(N, D, device=dev, dtype=torch.bfloat16), mods=torch.empty(N, W, **f),
            lmu=torch.empty(N, 8, **f), lsg=torch.empty(N, 8, **f), smu=torch.empty(N, **f), ssg=torch.empty(N, **f),
            xin=[torch.empty(M, D, **a) for _ in range(L + 1)], h1=[torch.empty(M, D, **a) for _ in range(L + 1)],
            qkv=[torch.empty(M, 3 * D, **a) for _ in range(L)], sc=[torch.empty(M, 2 * H, **f) for _ in range(L)],
            lse=[torch.empty(M, H, **f) for _ in range(L)], o=[torch.empty(M, D, **a) for _ in range(L)],
            a=[torch.empty(M, D, **a) for _ in range(L)], xmid=[torch.empty(M, D, **a) for _ in range(L)],
            h2=[torch.empty(M, D, **a) for _ in range(L)], z=[torch.empty(M, Hm, **a) for _ in range(L)],
            u=[torch.empty(M, Hm, **a) for _ in range(L)], b=[torch.empty(M, D, **a) for _ in range(L)],
            lin=torch.empty(M, ppc2, **a),
            # backward scratch
            R=torch.empty(M, D, **a), dY=torch.empty(M, D, **a), dY2=torch.empty(M, D, **a), dh=torch.empty(M, D, **a), dqkv=torch.empty(M, 3 * D, **a),
            dU=torch.empty(M, Hm, **a), dlin=torch.empty(M, ppc2, **a), dmods=torch.empty(N, W, **f),
            delta=torch.empty(M, H, **f), dgp=torch.empty(ops.modulate_bwd_partials(N, D), **f),
            dsmu=torch.empty(N, **f), dssg=torch.empty(N, **f), dlmu=torch.empty(N, 8, **f), dlsg=torch.empty(N, 8, **f),
            dc=torch.empty(N, D, **f), dcs=torch.empty(N, D, **f), dab=torch.empty(N, D, **f), dt1s=torch.empty(N, D, **f),
            P=torch.empty(M, K1, **f), R32=torch.empty(M, D, **f) if mode == "bf16" else None,
            dmods16=torch.empty(N, W, device=dev, dtype=torch.bfloat16) if mode == "bf16" else None,
            dWs=torch.empty(D * max(D, 256, K1), **f),  # d(effective weight) scratch of the small fp32 conditioning-path weights
        )
        if m.modulation != "adaln":  # (cos, sin) tables for the EPI_RESID_ROT epilogue (see Engine.workspace)
            B["rotcs"] = torch.empty(N, L * 2 * D, **f)
        if not m.flags["use_no_layernorm"]:  # LayerNorm statistics {mean, rstd} per token row, per modulate site
            B["ln1"] = [torch.empty(M, 2, **f) for _ in range(L + 1)]  # index L = final layer
            B["ln2"] = [torch.empty(M, 2, **f) for _ in range(L)]
        self._bufs[key] = B
        return B

    # ------------------------------------------------------------------ forward
    def forward(self, x, t, y, drop_mask):
        prev = ops.set_variant(self.m.variant)
        try:
            with torch.cuda.device(x.device):  # kernels launch on the current device's stream: make it the tensors' device
                out, saved = self._forward(x, t, y, drop_mask)
            self._generation += 1
            B = self.buffers(saved["N"], saved["mode"], saved["dev"])
            B["generation"] = saved["generation"] = self._generation
            return out, saved
        finally:
            ops.set_variant(prev)

    def backward(self, saved, dout):
        B = self.buffers(saved["N"], saved["mode"], saved["dev"])
        if B.get("generation") != saved.get("generation"):
            raise RuntimeError(
                "mapdit_b200: backward through a DiT forward whose saved activations were overwritten by a later train-mode "
                "forward of the same batch size (one forward may be outstanding per model and batch size; run forward -> "
                "backward pairs, or concatenate the inputs into one batch)")
        prev = ops.set_variant(self.m.variant)
        self._side_on = bool(self.wgrad_stream and saved["mode"] == "bf16")
        if self._side_on and (self._side is None or self._side.device != saved["dev"]):
            self._side = torch.cuda.Stream(device=saved["dev"])
        try:
            with torch.cuda.device(saved["dev"]):
                out = self._backward(saved, dout)
                self._join_side()  # every gradient is complete on the caller's stream
            return out
        finally:
            self._side_on = False
            self._side_reads.clear()
            self._wn_pending = []
            ops.set_variant(prev)

    def _forward(self, x, t, y, drop_mask):
        e, m = self.e, self.m
        mode = m.compute_dtype
        bf = mode == "bf16"
        N, dev = x.shape[0], x.device
        D, L, H = m.hidden_size, m.depth, m.num_heads
        hd = D // H
        T = (m.input_size // m.patch_size) ** 2
        x = x.contiguous().float()
        t = t.contiguous().to(torch.int64)
        y = y.contiguous().to(torch.int64)
        W = e.weights(mode, train=True)
        B = self.buffers(N, mode, dev)
        ld = B["mods"].shape[1]
        f = m.final_layer
        blk = m.blocks

        fl = m.flags
        ln, cosine = not fl["use_no_layernorm"], fl["use_cosine_attention"]
        if fl["use_mp_embedding"]:
            ops.fourier(t, m.t_embedder.embedding.scale, m.t_embedder.embedding.shift, B["e"])
        else:
            ops.timestep_sincos(t, B["e"])
        ops.gemm_f32(B["e"], W.wt1, out=B["t1"])
        ops.mp_silu(B["t1"], B["t1s"])
        ops.gemm_f32(B["t1s"], W.wt2, out=B["temb"])
        mask = None
        if m.y_embedder.dropout_prob > 0:
            if drop_mask is None:
                drop_mask = torch.rand(N, device=dev) < m.y_embedder.dropout_prob
            mask = drop_mask.to(torch.uint8).contiguous()
        ops.embed_rows(y, mask, m.num_classes, m.y_embedder.embedding.weight.data, B["yemb"])
        ops.cond_combine(B["temb"], B["yemb"], B["c"], B["cs"], B["cs16"])
        if bf:
            ops.gemm_bf16(B["cs16"], W.wmod, B["mods"])
        else:
            ops.gemm_f32(B["cs"], W.wmod, out=B["mods"])
        ops.gemm_f32(B["c"], W.wmu, out=B["lmu"])
        ops.gemm_f32(B["c"], W.wsg, out=B["lsg"])
        ops.mp_scale_from_lin(B["lmu"], f.mean_scale.reference.data, B["smu"])
        ops.mp_scale_from_lin(B["lsg"], f.sigma_scale.reference.data, B["ssg"])
        mods = B["mods"]

        lay = e.layout
        adaln = m.modulation == "adaln"
        fbase = L * lay["width"]

        def mod(i, name):
            return mods[:, i * lay["width"] + lay[name]:]

        def modulate_block(i, branch, src, dst):
            gain = (blk[i].gain_msa if branch == "a" else blk[i].gain_mlp).data
            if ln:
                ops.ln_modulate(src, dst, mod(i, "shift_" + branch), mod(i, "scale_" + branch),
                                B["ln1" if branch == "a" else "ln2"][i], ld, T)
            elif adaln:
                ops.modulate(src, dst, mod(i, "shift_" + branch), mod(i, "scale_" + branch), gain, ld, T)
            else:
                sc = mod(i, "scale_" + branch) if ("scale_" + branch) in lay else None
                ops.rotmod(src, dst, mod(i, "rot_" + branch), sc, gain, ld, T)

        def modulate_next(i, src, dst):
            if i + 1 < L:
                modulate_block(i + 1, "a", src, dst)
            elif ln:
                ops.ln_modulate(src, dst, mods[:, fbase:], mods[:, fbase + D:], B["ln1"][L], ld, T)
            else:
                ops.modulate(src, dst, mods[:, fbase:], mods[:, fbase + D:], f.gain_mod.data, ld, T)

        def qkv_proj(i, src, dst):
            if cosine and hd == 64:
                ops.gemm_bf16(src, W.wqkv[i], dst, epilogue=_lib.EPI_QKNORM, tokens=T, head_dim=hd, qk_cols=2 * D, aux=B["sc"][i])
            else:  # head_dim 72 (DiT-XL) or plain dot-product attention: plain store + standalone normalisation
                ops.gemm_bf16(src, W.wqkv[i], dst)
                if cosine:
                    ops.qk_normalize_save(dst, B["sc"][i], D, hd)

        fused = bf and adaln and not ln and cosine  # head_dim 72 (DiT-XL) only differs in qkv_proj
        fused_rot = bf and not adaln and not ln and cosine
        if fused_rot:
            rcs = B["rotcs"]
            for i in range(L):
                ops.rot_table(mod(i, "rot_a"), blk[i].gain_msa.data, rcs[:, (2 * i) * D:], ld, D,
                              mod(i, "rot_m"), blk[i].gain_mlp.data, rcs[:, (2 * i + 1) * D:])
            has_sc = "scale_a" in lay
        if adaln and not ln:
            ops.patch_embed(x, W.wx, m.pos_embed, B["xin"][0], B["h1"][0], mod(0, "shift_a"), mod(0, "scale_a"), blk[0].gain_msa.data,
                            ld, m.patch_size)
        else:
            ops.patch_embed(x, W.wx, m.pos_embed, B["xin"][0], None, None, None, None, ld, m.patch_size)
            modulate_block(0, "a", B["xin"][0], B["h1"][0])
        for i in range(L):
            xin, h1, qkv, o, a, xmid, h2, z, u, b = (B[k][i] for k in ("xin", "h1", "qkv", "o", "a", "xmid", "h2", "z", "u", "b"))
            xnext, hnext = B["xin"][i + 1], B["h1"][i + 1]
            if fused:
                if i + 1 < L:
                    nsh, nsc, ngn = mod(i + 1, "shift_a"), mod(i + 1, "scale_a"), blk[i + 1].gain_msa.data
                else:
                    nsh, nsc, ngn = mods[:, fbase:], mods[:, fbase + D:], f.gain_mod.data
                qkv_proj(i, h1, qkv)
                ops.cos_attn(qkv, o, N, T, H, hd, lse=B["lse"][i])
                ops.gemm_bf16(o, W.wo[i], xmid, epilogue=_lib.EPI_RESID_MOD, out2=h2, resid=xin, gate=mod(i, "gate_a"),
                              shift=mod(i, "shift_m"), scale=mod(i, "scale_m"), gain=blk[i].gain_mlp.data, ldmod=ld, tokens=T, aux=a)
                ops.gemm_bf16(h2, W.w1[i], u, epilogue=_lib.EPI_MPSILU, out2=z)
                ops.gemm_bf16(u, W.w2[i], xnext, epilogue=_lib.EPI_RESID_MOD, out2=hnext, resid=xmid, gate=mod(i, "gate_m"), shift=nsh,
                              scale=nsc, gain=ngn, ldmod=ld, tokens=T, aux=b)
            elif fused_rot:
                qkv_proj(i, h1, qkv)
                ops.cos_attn(qkv, o, N, T, H, hd, lse=B["lse"][i])
                ops.gemm_bf16(o, W.wo[i], xmid, epilogue=_lib.EPI_RESID_ROT, out2=h2, resid=xin, gate=mod(i, "gate_a"),
                              shift=rcs[:, (2 * i + 1) * D:], scale=mod(i, "scale_m") if has_sc else None, ldmod=ld,
                              ldrot=rcs.stride(0), tokens=T, aux=a)
                ops.gemm_bf16(h2, W.w1[i], u, epilogue=_lib.EPI_MPSILU, out2=z)
                if i + 1 < L:
                    ops.gemm_bf16(u, W.w2[i], xnext, epilogue=_lib.EPI_RESID_ROT, out2=hnext, resid=xmid, gate=mod(i, "gate_m"),
                                  shift=rcs[:, (2 * i + 2) * D:], scale=mod(i + 1, "scale_a") if has_sc else None, ldmod=ld,
                                  ldrot=rcs.stride(0), tokens=T, aux=b)
                else:
                    ops.gemm_bf16(u, W.w2[i], xnext, epilogue=_lib.EPI_RESID_MOD, out2=hnext, resid=xmid, gate=mod(i, "gate_m"),
                                  shift=mods[:, fbase:], scale=mods[:, fbase + D:], gain=f.gain_mod.data, ldmod=ld, tokens=T, aux=b)
            elif bf:
                qkv_proj(i, h1, qkv)
                ops.cos_attn(qkv, o, N, T, H, hd, lse=B["lse"][i])
                ops.gemm_bf16(o, W.wo[i], xmid, epilogue=_lib.EPI_RESID, resid=xin, gate=mod(i, "gate_a"), ldmod=ld, tokens=T, aux=a)
                modulate_block(i, "m", xmid, h2)
                ops.gemm_bf16(h2, W.w1[i], u, epilogue=_lib.EPI_MPSILU, out2=z)
                ops.gemm_bf16(u, W.w2[i], xnext, epilogue=_lib.EPI_RESID, resid=xmid, gate=mod(i, "gate_m"), ldmod=ld, tokens=T, aux=b)
                modulate_next(i, xnext, hnext)
            else:
                ops.gemm_f32(h1, W.wqkv[i], out=qkv)
                if cosine:
                    ops.qk_normalize_save(qkv, B["sc"][i], D, hd)
                ops.cos_attn(qkv, o, N, T, H, hd, lse=B["lse"][i])
                ops.gemm_f32(o, W.wo[i], out=a)
                ops.resid(xin, a, xmid, mod(i, "gate_a"), ld, T)
                modulate_block(i, "m", xmid, h2)
                ops.gemm_f32(h2, W.w1[i], out=z)
                ops.mp_silu(z, u)
                ops.gemm_f32(u, W.w2[i], out=b)
                ops.resid(xmid, b, xnext, mod(i, "gate_m"), ld, T)
                modulate_next(i, xnext, hnext)
        hF = B["h1"][L]
        if bf:
            ops.gemm_bf16(hF, W.wfl, B["lin"])
        else:
            ops.gemm_f32(hF, W.wfl, out=B["lin"])
        out = torch.empty(N, 2 * m.in_channels, m.input_size, m.input_size, device=dev, dtype=torch.float32)
        ops.final_unpatchify(B["lin"], B["smu"], B["ssg"], out, m.patch_size)
        saved = dict(N=N, mode=mode, dev=dev, x=x, t=t, y=y, mask=mask)
        return out, saved

    # ------------------------------------------------------------------ backward helpers
    def _dgrad(self, dy, w_eff, w_eff_t, out, bf):
        """out[M, K_in] = dy[M, N_out] @ W_eff[N_out, K_in]"""
        if bf:
            ops.gemm_bf16(dy, w_eff_t, out)
        else:
            ops.gemm_f32(dy, w_eff, out=out, trans_b=True)

    def _wgrad(self, dy, xin, param, bf, B, grads, reads=None):
        """param.grad = weight_norm_bwd(forced param, dy^T @ xin).  `reads` names the scratch buffer `dy` lives in: with the
        second stream active the caller must call _before_write(name) before the main stream overwrites that buffer."""
        g = self._gbuf(param)
        grads[id(param)] = g
        # the GEMM writes d(effective weight) straight into the gradient buffer; _flush_wn turns every pending buffer of a block
        # into d(raw weight) in ONE in-place multi-tensor launch (was a scratch buffer + one weight_norm_bwd launch per weight)
        batched = self.m.flags["use_weight_normalization"] and ops.WeightNormBwdBatch.supports(param.data, g)
        if self._side_on:
            main = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(main)
            self._side.wait_event(ev)  # dy is complete
            with torch.cuda.stream(self._side):
                ops.gemm_bf16_tn(dy, xin, g)
                if batched:
                    self._wn_pending.append((param, g))
                elif self.m.flags["use_weight_normalization"]:
                    ops.weight_norm_bwd(param.data, g.clone(), g)
                if reads is not None:
                    done = torch.cuda.Event()
                    done.record(self._side)
                    self._side_reads[reads] = done
            return
        if bf:
            ops.gemm_bf16_tn(dy, xin, g)
        else:
            ops.gemm_f32(dy, xin, out=g, trans_a=True, trans_b=True)
        if batched:
            self._wn_pending.append((param, g))
        elif self.m.flags["use_weight_normalization"]:
            ops.weight_norm_bwd(param.data, g.clone(), g)

    def _flush_wn(self):
        """weight-norm backward of every pending weight gradient, in place, one launch (on the stream the GEMMs ran on)"""
        if not self._wn_pending:
            return
        items = [(p.data, g) for p, g in self._wn_pending]
        self._wn_pending = []
        sig = tuple(t.data_ptr() for it in items for t in it)
        tab = self._wn_tables.get(sig)
        if tab is None:
            if len(self._wn_tables) > 256:  # autograd path: fresh gradient tensors every call
                self._wn_tables.clear()
            tab = self._wn_tables[sig] = ops.WeightNormBwdBatch(items, items[0][0].device)
        if self._side_on:
            with torch.cuda.stream(self._side):
                tab.run()
        else:
            tab.run()

    def _before_write(self, name):
        """main stream is about to overwrite scratch buffer `name`: wait for the side-stream GEMM that still reads it"""
        ev = self._side_reads.pop(name, None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def _join_side(self):
        if self._side_on:
            torch.cuda.current_stream().wait_stream(self._side)
            self._side_reads.clear()

    def _wn_bwd(self, param, dW, g=None):
        """gradient of the raw parameter from the gradient of its effective weight (tangent projection of the
        weight normalisation, or the identity with use_weight_normalization=False)"""
        if g is None:
            g = self._gbuf(param)
        if self.m.flags["use_weight_normalization"]:
            ops.weight_norm_bwd(param.data, dW, g)
        else:
            ops.axpby(dW, g, 1.0)
        return g

    def _gbuf(self, p):
        """gradient destination of parameter p: a view of the caller's flat gradient buffer, or a fresh tensor"""
        if self.grad_buffers is not None:
            return self.grad_buffers[id(p)]
        return torch.empty_like(p.data)

    def _scalar_from_partials(self, B, n, p):
        g = self._gbuf(p)
        ops.sum_partials(B["dgp"], n, g)
        return g

    # ------------------------------------------------------------------ backward
    def _backward(self, saved, dout):
        e, m = self.e, self.m
        N, mode, dev = saved["N"], saved["mode"], saved["dev"]
        bf = mode == "bf16"
        D, L, H = m.hidden_size, m.depth, m.num_heads
        hd = D // H
        T = (m.input_size // m.patch_size) ** 2
        W = e._w[mode]
        B = self.buffers(N, mode, dev)
        ld = B["mods"].shape[1]
        f = m.final_layer
        blk = m.blocks
        mods, dmods = B["mods"], B["dmods"]
        npart = ops.modulate_bwd_partials(N, D) if m.modulation == "adaln" else ops.rotmod_bwd_partials(N, D)
        grads = {}

        lay = e.layout
        adaln = m.modulation == "adaln"
        fbase = L * lay["width"]

        def mod(buf, i, name):
            return buf[:, i * lay["width"] + lay[name]:]

        fl = m.flags
        ln, cosine = not fl["use_no_layernorm"], fl["use_cosine_attention"]

        def zero_gain_grad(gp):  # LayerNorm adaLN does not use the gain parameter
            g = self._gbuf(gp)
            g.zero_()
            grads[id(gp)] = g

        fuse_resid = not ln  # the residual backward that follows a modulate / rotation backward runs in the same kernel

        def resid_of(i, branch):
            """(y, gate, dgate, dy) of block i's residual after branch 'a' / 'm'; the two branches use separate dy buffers so the
            side-stream weight-gradient GEMM of one can still read its dy while the main stream produces the other's"""
            self._before_write("dYa" if branch == "a" else "dYm")
            return (B["a" if branch == "a" else "b"][i], mod(mods, i, "gate_" + branch), mod(dmods, i, "gate_" + branch),
                    dYa if branch == "a" else dYm)

        def modulate_block_bwd(i, branch, dh_, x_, accumulate, then_resid=None):
            """backward of the block-i modulation: R (+)= d/dx, per-sample vector grads into dmods, returns d(gain).
            `then_resid` = (i', branch') also applies the backward of that residual to the updated R (fused kernel)."""
            gp = blk[i].gain_msa if branch == "a" else blk[i].gain_mlp
            if then_resid is not None:
                y_, gate_, dgate_, dY_ = resid_of(*then_resid)
                if adaln:
                    ops.modulate_resid_bwd(dh_, x_, R, mod(mods, i, "shift_" + branch), mod(mods, i, "scale_" + branch), gp.data,
                                           mod(dmods, i, "shift_" + branch), mod(dmods, i, "scale_" + branch), B["dgp"], y_, dY_, gate_,
                                           dgate_, ld, N, T, accumulate)
                else:
                    has_sc = ("scale_" + branch) in lay
                    ops.rotmod_resid_bwd(dh_, x_, R, mod(mods, i, "rot_" + branch), mod(mods, i, "scale_" + branch) if has_sc else None,
                                         gp.data, mod(dmods, i, "rot_" + branch), mod(dmods, i, "scale_" + branch) if has_sc else None,
                                         B["dgp"], y_, dY_, gate_, dgate_, ld, N, T, accumulate)
                grads[id(gp)] = self._scalar_from_partials(B, npart, gp)
                return
            if ln:
                ops.ln_modulate_bwd(dh_, x_, R, B["ln1" if branch == "a" else "ln2"][i], mod(mods, i, "scale_" + branch),
                                    mod(dmods, i, "shift_" + branch), mod(dmods, i, "scale_" + branch), ld, N, T, accumulate)
                zero_gain_grad(gp)
                return
            if adaln:
                ops.modulate_bwd(dh_, x_, R, mod(mods, i, "shift_" + branch), mod(mods, i, "scale_" + branch), gp.data,
                                 mod(dmods, i, "shift_" + branch), mod(dmods, i, "scale_" + branch), B["dgp"], ld, N, T, accumulate)
            else:
                has_sc = ("scale_" + branch) in lay
                ops.rotmod_bwd(dh_, x_, R, mod(mods, i, "rot_" + branch), mod(mods, i, "scale_" + branch) if has_sc else None, gp.data,
                               mod(dmods, i, "rot_" + branch), mod(dmods, i, "scale_" + branch) if has_sc else None, B["dgp"], ld, N, T,
                               accumulate)
            grads[id(gp)] = self._scalar_from_partials(B, npart, gp)

        R, dYm, dYa, dh, dqkv, dU = B["R"], B["dY"], B["dY2"], B["dh"], B["dqkv"], B["dU"]
        # ---- final layer (src/blocks/final_layer.py:53-59)
        ops.final_bwd(dout, B["lin"], B["smu"], B["ssg"], B["dlin"], B["dsmu"], B["dssg"], m.patch_size)
        gref_mu, gref_sg = self._gbuf(f.mean_scale.reference), self._gbuf(f.sigma_scale.reference)
        ops.mp_scale_bwd(B["dsmu"], B["smu"], B["lmu"], f.mean_scale.reference.data, B["dlmu"], gref_mu)
        ops.mp_scale_bwd(B["dssg"], B["ssg"], B["lsg"], f.sigma_scale.reference.data, B["dlsg"], gref_sg)
        grads[id(f.mean_scale.reference)] = gref_mu
        grads[id(f.sigma_scale.reference)] = gref_sg
        ops.gemm_f32(B["dlmu"], W.wmu, out=B["dc"], trans_b=True)                      # dc = dl_mu @ W_mu
        ops.gemm_f32(B["dlsg"], W.wsg, out=B["dc"], trans_b=True, accumulate=True)
        for dl, p in ((B["dlmu"], f.mean_scale.linear.weight), (B["dlsg"], f.sigma_scale.linear.weight)):
            dW = B["dWs"][: 8 * D].view(8, D)
            ops.gemm_f32(dl, B["c"], out=dW, trans_a=True, trans_b=True)
            grads[id(p)] = self._wn_bwd(p, dW)
        hF, xF = B["h1"][L], B["xin"][L]
        self._wgrad(B["dlin"], hF, f.linear.weight, bf, B, grads)
        self._flush_wn()
        self._dgrad(B["dlin"], W.wfl, getattr(W, "wfl_t", None), dh, bf)
        if ln:
            ops.ln_modulate_bwd(dh, xF, R, B["ln1"][L], mods[:, fbase + D:], dmods[:, fbase:], dmods[:, fbase + D:], ld, N, T, False)
            zero_gain_grad(f.gain_mod)
        else:
            if fuse_resid:  # + the backward of the last block's MLP residual
                y_, gate_, dgate_, dY_ = resid_of(L - 1, "m")
                ops.modulate_resid_bwd(dh, xF, R, mods[:, fbase:], mods[:, fbase + D:], f.gain_mod.data, dmods[:, fbase:],
                                       dmods[:, fbase + D:], B["dgp"], y_, dY_, gate_, dgate_, ld, N, T, False)
            else:
                ops.modulate_bwd(dh, xF, R, mods[:, fbase:], mods[:, fbase + D:], f.gain_mod.data, dmods[:, fbase:],
                                 dmods[:, fbase + D:], B["dgp"], ld, N, T, False)
            grads[id(f.gain_mod)] = self._scalar_from_partials(B, ops.modulate_bwd_partials(N, D), f.gain_mod)
        # ---- blocks, last to first
        for i in range(L - 1, -1, -1):
            xin, h1, qkv, o, a, xmid, h2, z, u, b = (B[k][i] for k in ("xin", "h1", "qkv", "o", "a", "xmid", "h2", "z", "u", "b"))
            wt = (lambda name: getattr(W, name)[i]) if bf else (lambda name: None)
            # MLP branch
            if not fuse_resid:
                self._before_write("dYm")
                ops.resid_bwd(R, b, dYm, mod(mods, i, "gate_m"), mod(dmods, i, "gate_m"), ld, N, T)
            self._wgrad(dYm, u, blk[i].mlp.net[2].weight, bf, B, grads, reads="dYm")
            self._before_write("dU")
            if bf:  # dgrad of fc2 with MPSiLU's backward fused into the epilogue
                ops.gemm_bf16(dYm, W.w2_t[i], dU, epilogue=_lib.EPI_SILU_BWD, resid=z)
            else:
                self._dgrad(dYm, W.w2[i], None, dU, bf)
                ops.mp_silu_bwd(dU, z, dU)
            self._wgrad(dU, h2, blk[i].mlp.net[0].weight, bf, B, grads, reads="dU")
            self._dgrad(dU, W.w1[i], wt("w1_t"), dh, bf)
            modulate_block_bwd(i, "m", dh, xmid, True, then_resid=(i, "a") if fuse_resid else None)
            # attention branch
            if not fuse_resid:
                self._before_write("dYa")
                ops.resid_bwd(R, a, dYa, mod(mods, i, "gate_a"), mod(dmods, i, "gate_a"), ld, N, T)
            self._wgrad(dYa, o, blk[i].attn.out_proj.weight, bf, B, grads, reads="dYa")
            # bf16, 256 tokens, head_dim 64: the out-proj dgrad GEMM also emits delta = dO.O per (row, head) in its epilogue, so the
            # fused attention backward does not need a separate pass over dO and O
            fuse_delta = bf and cosine and hd == 64 and T == 256 and self.fuse_delta
            if fuse_delta:
                ops.gemm_bf16(dYa, W.wo_t[i], dh, epilogue=_lib.EPI_STORE_DELTA, resid=o, aux=B["delta"])
            else:
                self._dgrad(dYa, W.wo[i], wt("wo_t"), dh, bf)
            self._before_write("dqkv")
            if cosine:  # attention backward with the q/k normalisation backward fused into its dq / dk epilogues
                ops.cos_attn_bwd_qknorm(qkv, None if fuse_delta else o, dh, B["lse"][i], B["sc"][i], dqkv, B["delta"], N, T, H, hd)
            else:
                ops.cos_attn_bwd(qkv, o, dh, B["lse"][i], dqkv, B["delta"], N, T, H, hd)
            self._wgrad(dqkv, h1, blk[i].attn.qkv_proj.weight, bf, B, grads, reads="dqkv")
            self._dgrad(dqkv, W.wqkv[i], wt("wqkv_t"), dh, bf)
            modulate_block_bwd(i, "a", dh, xin, True, then_resid=(i - 1, "m") if (fuse_resid and i > 0) else None)
            # block i's modulation vectors have their final gradient: its modulation-weight gradient now (not in one GEMM over
            # all blocks after the loop), so it rides in the block's all-reduce bucket and overlaps the rest of the backward
            sl = slice(i * lay["width"], (i + 1) * lay["width"])
            if bf:
                ops.cast_2d(dmods[:, sl], B["dmods16"][:, sl])
                self._wgrad(B["dmods16"][:, sl], B["cs16"], blk[i].modulation[1].weight, bf, B, grads)
            else:
                self._wgrad(dmods[:, sl], B["cs"], blk[i].modulation[1].weight, bf, B, grads)
            self._flush_wn()
            if self.grad_hook is not None:
                pairs = [(p, grads[id(p)]) for p in blk[i].parameters() if id(p) in grads]
                if self._side_on:  # the block's weight gradients complete on the second stream: fire the hook (NCCL) there
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream())
                    self._side.wait_event(ev)  # ... after the gain gradients the main stream just produced
                    with torch.cuda.stream(self._side):
                        self.grad_hook(pairs)
                else:
                    self.grad_hook(pairs)
        self._join_side()  # the tail uses the dWs scratch on the main stream
        # ---- patch embed (src/dit.py:81-84): x0 = (lin + pos)/2/sqrt(.5) -> d lin = R * 0.5/sqrt(.5)
        K1 = m.x_embedder.weight.shape[1]
        dWx = B["dWs"][: D * K1].view(D, K1)
        ops.patch_embed_wgrad(R, saved["x"], dWx, m.patch_size, 0.5 / 0.7071067811865476 if fl["use_mp_pos_enc"] else 1.0)
        grads[id(m.x_embedder.weight)] = self._wn_bwd(m.x_embedder.weight, dWx)
        # ---- modulation GEMM: one dgrad over the concatenated weight of all blocks + final layer (the blocks' weight gradients
        # were taken inside the loop), the final layer's weight gradient
        if bf:
            ops.cast_2d(dmods[:, fbase:], B["dmods16"][:, fbase:])
            ops.gemm_bf16(B["dmods16"], W.wmod_t, B["dcs"])
            self._wgrad(B["dmods16"][:, fbase:], B["cs16"], f.modulation[1].weight, bf, B, grads)
        else:
            ops.gemm_f32(dmods, W.wmod, out=B["dcs"], trans_b=True)
            self._wgrad(dmods[:, fbase:], B["cs"], f.modulation[1].weight, bf, B, grads)
        self._flush_wn()
        # ---- conditioning path (src/dit.py:86-88)
        ops.cond_combine_bwd(B["c"], B["dc"], B["dcs"], B["dab"])
        table = m.y_embedder.embedding.weight
        gt = self._gbuf(table)
        gt.zero_()
        ops.embed_rows_bwd(saved["y"], saved["mask"], m.num_classes, table.data, B["dab"], gt)
        grads[id(table)] = gt
        p2, p1 = m.t_embedder.mlp.net[2].weight, m.t_embedder.mlp.net[0].weight
        dW = B["dWs"][: D * D].view(D, D)
        ops.gemm_f32(B["dab"], B["t1s"], out=dW, trans_a=True, trans_b=True)
        grads[id(p2)] = self._wn_bwd(p2, dW)
        ops.gemm_f32(B["dab"], W.wt2, out=B["dt1s"], trans_b=True)
        ops.mp_silu_bwd(B["dt1s"], B["t1"], B["dt1s"])
        dW = B["dWs"][: D * 256].view(D, 256)
        ops.gemm_f32(B["dt1s"], B["e"], out=dW, trans_a=True, trans_b=True)
        grads[id(p1)] = self._wn_bwd(p1, dW)
        self._join_side()  # the final layer's modulation-weight gradient ran on the second stream
        if self.grad_hook is not None:
            done = {id(q) for b_ in blk for q in b_.parameters()}
            self.grad_hook([(q, grads[id(q)]) for q in m.parameters() if id(q) not in done])
        return [grads[id(p)] for p in m.parameters()]
