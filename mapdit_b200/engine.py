"""Kernel schedule of one DiT forward (and, in train mode, the matching backward).

HBM layout: token-major activations [M = N*T, features] in the activation dtype (bf16 on the
tcgen05 path, fp32 in parity mode); per-sample conditioning [N, *] always fp32; effective
(normalised) weights cached per parameter version, the modulation weights of all blocks plus the
final layer concatenated so one GEMM produces every shift/scale/gate of the step.

Reference order of operations: src/dit.py:70-105 -> src/blocks/dit_block.py:32-37 ->
src/layers/attention.py:27-51, src/layers/mlp.py:18-25 -> src/blocks/final_layer.py:53-59.
"""
import torch

from . import _lib, ops


class _Weights:
    """effective weights W_row/(||W_row||+eps) (src/basic/mp_linear.py:42-46) in the layouts the kernels read."""
    pass


class Engine:
    def __init__(self, model):
        self.m = model
        self._w = {}       # mode -> _Weights
        self._w_sig = {}   # mode -> signature of parameter versions
        self._ws = {}      # (N, mode, device) -> workspace dict
        self._dirty = True

    # ------------------------------------------------------------------ parameters
    def _linear_params(self):
        m = self.m
        ps = [m.x_embedder.weight, m.t_embedder.mlp.net[0].weight, m.t_embedder.mlp.net[2].weight,
              m.y_embedder.embedding.weight]
        for b in m.blocks:
            ps += [b.attn.qkv_proj.weight, b.attn.out_proj.weight, b.mlp.net[0].weight, b.mlp.net[2].weight,
                   b.modulation[1].weight]
        f = m.final_layer
        ps += [f.linear.weight, f.modulation[1].weight, f.mean_scale.linear.weight, f.sigma_scale.linear.weight]
        return ps

    def _signature(self):
        ps = self._linear_params()
        return (sum(p._version for p in ps), tuple(p.data_ptr() for p in ps[:3]), ps[0].device)

    def invalidate(self):
        self._dirty = True

    @property
    def layout(self):
        from .dit import mod_layout
        return mod_layout(self.m.modulation, self.m.hidden_size)

    @property
    def trainer(self):
        if getattr(self, "_trainer", None) is None:
            from .autograd import Trainer
            self._trainer = Trainer(self)
        return self._trainer

    def weights(self, mode, train):
        """(Re)build the effective weights.  Train mode always re-normalises, writing the forced
        normalisation back into the parameters first (src/basic/mp_linear.py:37-40)."""
        sig = self._signature()
        if not train and not self._dirty and self._w_sig.get(mode) == sig and mode in self._w:
            return self._w[mode]
        m = self.m
        dev = m.x_embedder.weight.device
        D, L = m.hidden_size, m.depth
        wdt = torch.bfloat16 if mode == "bf16" else torch.float32
        W = self._w.get(mode)
        if W is None or W.device != dev:
            W = _Weights()
            W.device = dev
            f32 = dict(device=dev, dtype=torch.float32)
            wd = dict(device=dev, dtype=wdt)
            W.wx = torch.empty_like(m.x_embedder.weight, **f32)
            W.wt1 = torch.empty(D, 256, **f32)
            W.wt2 = torch.empty(D, D, **f32)
            W.wmu = torch.empty(8, D, **f32)
            W.wsg = torch.empty(8, D, **f32)
            W.modw = self.layout["width"]
            W.mod_total = L * W.modw + 2 * D
            W.wmod = torch.empty(W.mod_total, D, **wd)
            Hm = m.blocks[0].mlp.hidden_dim
            W.wqkv = [torch.empty(3 * D, D, **wd) for _ in range(L)]
            W.wo = [torch.empty(D, D, **wd) for _ in range(L)]
            W.w1 = [torch.empty(Hm, D, **wd) for _ in range(L)]
            W.w2 = [torch.empty(D, Hm, **wd) for _ in range(L)]
            W.wfl = torch.empty_like(m.final_layer.linear.weight, **wd)
            self._w[mode] = W
        fl = m.flags
        # 1 = forced write-back + normalise (train), 0 = normalise only, -1 = raw weights (use_weight_normalization=False)
        force = -1 if not fl["use_weight_normalization"] else int(bool(train) and fl["use_forced_weight_normalization"])
        want_t = bool(train) and mode == "bf16"  # transposed bf16 copies feed the dgrad GEMMs
        if want_t and not hasattr(W, "wqkv_t"):
            wd = dict(device=dev, dtype=wdt)
            Hm = m.blocks[0].mlp.hidden_dim
            W.wqkv_t = [torch.empty(D, 3 * D, **wd) for _ in range(L)]
            W.wo_t = [torch.empty(D, D, **wd) for _ in range(L)]
            W.w1_t = [torch.empty(D, Hm, **wd) for _ in range(L)]
            W.w2_t = [torch.empty(Hm, D, **wd) for _ in range(L)]
            W.wfl_t = torch.empty(D, m.final_layer.linear.weight.shape[0], **wd)
            W.wmod_t = torch.empty(D, W.mod_total, **wd)

        batch = []  # big weights go through ONE multi-tensor launch pair (ops.WeightNormBatch)

        def norm(p, out, out_t=None, ld_t=0):
            kw = {"eff_bf16": out} if out.dtype == torch.bfloat16 else {"eff_f32": out}
            if want_t and out_t is not None:
                kw["eff_bf16_t"] = out_t
                kw["ld_t"] = ld_t
            w = p.data
            ok = (w.shape[1] % 4 == 0 and w.shape[0] >= 16 and w.is_contiguous() and w.data_ptr() % 16 == 0
                  and out.data_ptr() % 16 == 0 and (kw.get("eff_bf16_t") is None or kw["eff_bf16_t"].data_ptr() % 2 == 0))
            if ok:
                batch.append((w, kw.get("eff_f32"), kw.get("eff_bf16"), kw.get("eff_bf16_t"), kw.get("ld_t", 0)))
            else:
                ops.weight_norm_fwd(w, force=force, **kw)

        norm(m.x_embedder.weight, W.wx)
        norm(m.t_embedder.mlp.net[0].weight, W.wt1)
        norm(m.t_embedder.mlp.net[2].weight, W.wt2)
        if force > 0 and fl["use_mp_embedding"]:  # the embedding table is normalised in place too (src/basic/mp_embedding.py:16-19)
            ops.weight_norm_fwd(m.y_embedder.embedding.weight.data, force=True)
        tget = (lambda name, i: getattr(W, name)[i]) if want_t else (lambda name, i: None)
        for i, b in enumerate(m.blocks):
            norm(b.attn.qkv_proj.weight, W.wqkv[i], tget("wqkv_t", i))
            norm(b.attn.out_proj.weight, W.wo[i], tget("wo_t", i))
            norm(b.mlp.net[0].weight, W.w1[i], tget("w1_t", i))
            norm(b.mlp.net[2].weight, W.w2[i], tget("w2_t", i))
            norm(b.modulation[1].weight, W.wmod[i * W.modw:(i + 1) * W.modw], W.wmod_t[:, i * W.modw:] if want_t else None, W.mod_total)
        f = m.final_layer
        norm(f.modulation[1].weight, W.wmod[L * W.modw:], W.wmod_t[:, L * W.modw:] if want_t else None, W.mod_total)
        norm(f.linear.weight, W.wfl, W.wfl_t if want_t else None)
        norm(f.mean_scale.linear.weight, W.wmu)
        norm(f.sigma_scale.linear.weight, W.wsg)
        if batch:
            key = (mode, want_t)
            sigb = tuple(t.data_ptr() for it in batch for t in it[:4] if t is not None)
            wb = getattr(self, "_wn_batches", {}).get(key)
            if wb is None or wb[0] != sigb:
                wb = (sigb, ops.WeightNormBatch(batch, dev))
                self.__dict__.setdefault("_wn_batches", {})[key] = wb
            wb[1].run(force)
        self._w_sig[mode] = self._signature()
        self._dirty = bool(train)  # forced WN rewrote the parameters without bumping versions
        return W

    # ------------------------------------------------------------------ workspaces
    def workspace(self, N, mode, dev):
        key = (N, mode, str(dev))
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        m = self.m
        D, T = m.hidden_size, (m.input_size // m.patch_size) ** 2
        M = N * T
        Hm = m.blocks[0].mlp.hidden_dim
        adt = torch.bfloat16 if mode == "bf16" else torch.float32
        a = dict(device=dev, dtype=adt)
        f = dict(device=dev, dtype=torch.float32)
        ppc2 = m.final_layer.linear.weight.shape[0]
        ws = dict(
            x=torch.empty(M, D, **a), h=torch.empty(M, D, **a), qkv=torch.empty(M, 3 * D, **a), o=torch.empty(M, D, **a),
            u=torch.empty(M, Hm, **a), lin=torch.empty(M, ppc2, **a),
            e=torch.empty(N, 256, **f), t1=torch.empty(N, D, **f), t1s=torch.empty(N, D, **f), temb=torch.empty(N, D, **f),
            yemb=torch.empty(N, D, **f), c=torch.empty(N, D, **f), cs=torch.empty(N, D, **f),
            cs16=torch.empty(N, D, device=dev, dtype=torch.bfloat16),
            mods=torch.empty(N, m.depth * self.layout["width"] + 2 * D, **f), smu=torch.empty(N, **f), ssg=torch.empty(N, **f),
        )
        if mode == "fp32":
            ws["tmp"] = torch.empty(M, D, **a)
        if m.modulation != "adaln":  # (cos, sin) tables of every block's two rotations, read by the EPI_RESID_ROT epilogue
            ws["rotcs"] = torch.empty(N, m.depth * 2 * D, **f)
        self._ws[key] = ws
        return ws

    _COND_KEYS = ("mods", "smu", "ssg", "rotcs")

    def cond_slot(self, ws, slot):
        """the conditioning outputs (modulation vectors, rotation tables, MPScale factors) of `ws`, slot 0 = the workspace's own
        buffers, slot 1 = a second set (allocated on first use): the graphed sampling loop computes step k+1's conditioning into
        one slot on a side branch while step k's blocks read the other"""
        if slot == 0:
            return {k: ws[k] for k in self._COND_KEYS if k in ws}
        if "cond1" not in ws:
            ws["cond1"] = {k: torch.empty_like(ws[k]) for k in self._COND_KEYS if k in ws}
        return ws["cond1"]

    # ------------------------------------------------------------------ forward
    def forward(self, x, t, y, train=False, drop_mask=None, cond_ready=None):
        """`cond_ready` (eval only): conditioning slot that already holds the outputs of `conditioning(t, y)` — the chain at the head
        of the forward is skipped and the blocks read that slot"""
        m = self.m
        if not x.is_cuda:
            raise RuntimeError("mapdit_b200.DiT runs on CUDA only (hand-written sm_100a kernels, no CPU fallback)")
        mode = m.compute_dtype
        if torch.is_grad_enabled() and any(p.requires_grad for p in m.parameters()) and train:
            from .autograd import dit_forward_autograd
            return dit_forward_autograd(self, x, t, y, drop_mask)
        with torch.cuda.device(x.device):  # kernels launch on the current device's stream: make it the tensors' device
            return self._forward_impl(x, t, y, train, drop_mask, mode, save=None, cond_ready=cond_ready)

    def _forward_impl(self, x, t, y, train, drop_mask, mode, save, cond_ready=None):
        prev = ops.set_variant(self.m.variant)
        try:
            return self._forward_body(x, t, y, train, drop_mask, mode, save, cond_ready)
        finally:
            ops.set_variant(prev)

    def conditioning(self, t, y, slot=0, train=False, drop_mask=None):
        """the part of the forward that depends on (t, y) only (src/dit.py:86-88 + every block's modulation linear,
        src/blocks/dit_block.py:33, + MPScale of the final layer) into conditioning slot `slot`"""
        m = self.m
        mode = m.compute_dtype
        prev = ops.set_variant(m.variant)
        try:
            with torch.cuda.device(t.device):
                W = self.weights(mode, train)
                ws = self.workspace(t.shape[0], mode, t.device)
                self._conditioning(ws, W, t.contiguous().to(torch.int64), y.contiguous().to(torch.int64), train, drop_mask, mode,
                                   self.cond_slot(ws, slot))
        finally:
            ops.set_variant(prev)

    def _conditioning(self, ws, W, t, y, train, drop_mask, mode, cb):
        m = self.m
        N, dev = t.shape[0], t.device
        D, L = m.hidden_size, m.depth
        bf = mode == "bf16"
        ld = cb["mods"].shape[1]
        fl = m.flags
        if fl["use_mp_embedding"]:
            ops.fourier(t, m.t_embedder.embedding.scale, m.t_embedder.embedding.shift, ws["e"])
        else:
            ops.timestep_sincos(t, ws["e"])
        ops.gemm_f32(ws["e"], W.wt1, out=ws["t1"])
        ops.mp_silu(ws["t1"], ws["t1s"])
        ops.gemm_f32(ws["t1s"], W.wt2, out=ws["temb"])
        mask = None
        if train and m.y_embedder.dropout_prob > 0:
            if drop_mask is None:
                drop_mask = torch.rand(N, device=dev) < m.y_embedder.dropout_prob
            mask = drop_mask.to(torch.uint8).contiguous()
        ops.embed_rows(y, mask, m.num_classes, m.y_embedder.embedding.weight.data, ws["yemb"])
        ops.cond_combine(ws["temb"], ws["yemb"], ws["c"], ws["cs"], ws["cs16"])
        if bf:
            ops.gemm_bf16(ws["cs16"], W.wmod, cb["mods"])
        else:
            ops.gemm_f32(ws["cs"], W.wmod, out=cb["mods"])
        f = m.final_layer
        ops.mp_scale(ws["c"], W.wmu, f.mean_scale.reference.data, cb["smu"])
        ops.mp_scale(ws["c"], W.wsg, f.sigma_scale.reference.data, cb["ssg"])
        if "rotcs" in cb and self._fused_rot(mode):
            lay, mods, rcs, blk = self.layout, cb["mods"], cb["rotcs"], m.blocks
            for i in range(L):
                ops.rot_table(mods[:, i * lay["width"] + lay["rot_a"]:], blk[i].gain_msa.data, rcs[:, (2 * i) * D:], ld, D,
                              mods[:, i * lay["width"] + lay["rot_m"]:], blk[i].gain_mlp.data, rcs[:, (2 * i + 1) * D:])

    def _fused_rot(self, mode):
        fl = self.m.flags
        return mode == "bf16" and self.m.modulation != "adaln" and fl["use_no_layernorm"] and fl["use_cosine_attention"]

    def _forward_body(self, x, t, y, train, drop_mask, mode, save, cond_ready=None):
        m = self.m
        N = x.shape[0]
        dev = x.device
        D, L, T = m.hidden_size, m.depth, (m.input_size // m.patch_size) ** 2
        H, hd = m.num_heads, m.hidden_size // m.num_heads
        x = x.contiguous().float()
        W = self.weights(mode, train)
        ws = self.workspace(N, mode, dev)
        bf = mode == "bf16"
        fl = m.flags
        # ---- conditioning (fp32; src/dit.py:86-88, timestep_embedder.py:18-43, label_embedder.py:29-34)
        cb = self.cond_slot(ws, cond_ready or 0)
        if cond_ready is None:
            self._conditioning(ws, W, t.contiguous().to(torch.int64), y.contiguous().to(torch.int64), train, drop_mask, mode, cb)
        ld = cb["mods"].shape[1]
        f = m.final_layer
        mods = cb["mods"]
        return self._blocks(x, ws, W, cb, mods, ld, N, dev, D, L, T, H, hd, bf, fl, f, mode)

    def _blocks(self, x, ws, W, cb, mods, ld, N, dev, D, L, T, H, hd, bf, fl, f, mode):
        m = self.m
        lay = self.layout
        adaln = m.modulation == "adaln"
        ln = not fl["use_no_layernorm"]  # vanilla adaLN: LayerNorm then x(1+scale)+shift (UNPINNED)
        cosine = fl["use_cosine_attention"]
        fbase = L * lay["width"]  # final layer: [shift | scale]

        def mod(i, name):  # column slice `name` of block i's modulation output
            return mods[:, i * lay["width"] + lay[name]:]

        def modulate_block(i, branch, src, dst):
            """h = block-i modulation of the residual stream for branch 'a' (attention) / 'm' (MLP), standalone kernel"""
            gain = (blk[i].gain_msa if branch == "a" else blk[i].gain_mlp).data
            if ln:
                ops.ln_modulate(src, dst, mod(i, "shift_" + branch), mod(i, "scale_" + branch), None, ld, T)
            elif adaln:
                ops.modulate(src, dst, mod(i, "shift_" + branch), mod(i, "scale_" + branch), gain, ld, T)
            else:
                sc = mod(i, "scale_" + branch) if ("scale_" + branch) in lay else None
                ops.rotmod(src, dst, mod(i, "rot_" + branch), sc, gain, ld, T)

        def modulate_next(i, src, dst):
            if i + 1 < L:
                modulate_block(i + 1, "a", src, dst)
            elif ln:
                ops.ln_modulate(src, dst, mods[:, fbase:], mods[:, fbase + D:], None, ld, T)
            else:
                ops.modulate(src, dst, mods[:, fbase:], mods[:, fbase + D:], f.gain_mod.data, ld, T)

        def qkv_proj(i, src):
            """qkv GEMM (+ q/k L2 normalisation: fused epilogue for head_dim 64, standalone kernel otherwise)"""
            if cosine and hd == 64:
                ops.gemm_bf16(src, W.wqkv[i], ws["qkv"], epilogue=_lib.EPI_QKNORM, tokens=T, head_dim=hd, qk_cols=2 * D)
            else:
                ops.gemm_bf16(src, W.wqkv[i], ws["qkv"])
                if cosine:
                    ops.qk_normalize(ws["qkv"], D, hd)

        blk = m.blocks
        # modulate fused into the residual GEMM epilogues (the pinned all-flags-on configuration; head_dim 72 of DiT-XL only
        # differs in the qkv GEMM: plain store + the row-wise q/k normalisation kernel, see qkv_proj)
        fused = bf and adaln and not ln and cosine
        # the same for rotation(+scaling) modulation: BASELINE.json's headline configuration (UNPINNED)
        fused_rot = bf and not adaln and not ln and cosine
        if fused_rot:
            rcs = cb["rotcs"]  # filled by _conditioning
            has_sc = "scale_a" in lay
        # ---- patch embed + first modulate (src/dit.py:81-84)
        X, Hb = ws["x"], ws["h"]
        if adaln and not ln:
            ops.patch_embed(x, W.wx, m.pos_embed, X, Hb, mod(0, "shift_a"), mod(0, "scale_a"), blk[0].gain_msa.data, ld, m.patch_size)
        else:
            ops.patch_embed(x, W.wx, m.pos_embed, X, None, None, None, None, ld, m.patch_size)
            modulate_block(0, "a", X, Hb)
        for i in range(L):
            if fused:
                nxt_shift, nxt_scale, nxt_gain = ((mod(i + 1, "shift_a"), mod(i + 1, "scale_a"), blk[i + 1].gain_msa.data) if i + 1 < L
                                                  else (mods[:, fbase:], mods[:, fbase + D:], f.gain_mod.data))
                qkv_proj(i, Hb)
                ops.cos_attn(ws["qkv"], ws["o"], N, T, H, hd)
                ops.gemm_bf16(ws["o"], W.wo[i], X, epilogue=_lib.EPI_RESID_MOD, out2=Hb, resid=X, gate=mod(i, "gate_a"),
                              shift=mod(i, "shift_m"), scale=mod(i, "scale_m"), gain=blk[i].gain_mlp.data, ldmod=ld, tokens=T)
                ops.gemm_bf16(Hb, W.w1[i], ws["u"], epilogue=_lib.EPI_MPSILU)
                ops.gemm_bf16(ws["u"], W.w2[i], X, epilogue=_lib.EPI_RESID_MOD, out2=Hb, resid=X, gate=mod(i, "gate_m"),
                              shift=nxt_shift, scale=nxt_scale, gain=nxt_gain, ldmod=ld, tokens=T)
            elif fused_rot:
                qkv_proj(i, Hb)
                ops.cos_attn(ws["qkv"], ws["o"], N, T, H, hd)
                ops.gemm_bf16(ws["o"], W.wo[i], X, epilogue=_lib.EPI_RESID_ROT, out2=Hb, resid=X, gate=mod(i, "gate_a"),
                              shift=rcs[:, (2 * i + 1) * D:], scale=mod(i, "scale_m") if has_sc else None, ldmod=ld,
                              ldrot=rcs.stride(0), tokens=T)
                ops.gemm_bf16(Hb, W.w1[i], ws["u"], epilogue=_lib.EPI_MPSILU)
                if i + 1 < L:
                    ops.gemm_bf16(ws["u"], W.w2[i], X, epilogue=_lib.EPI_RESID_ROT, out2=Hb, resid=X, gate=mod(i, "gate_m"),
                                  shift=rcs[:, (2 * i + 2) * D:], scale=mod(i + 1, "scale_a") if has_sc else None, ldmod=ld,
                                  ldrot=rcs.stride(0), tokens=T)
                else:  # the final layer keeps the MP-AdaLN modulate (src/blocks/final_layer.py:53-59)
                    ops.gemm_bf16(ws["u"], W.w2[i], X, epilogue=_lib.EPI_RESID_MOD, out2=Hb, resid=X, gate=mod(i, "gate_m"),
                                  shift=mods[:, fbase:], scale=mods[:, fbase + D:], gain=f.gain_mod.data, ldmod=ld, tokens=T)
            elif bf:  # LayerNorm modulation, plain attention, head_dim != 64: residual fused, rest standalone
                qkv_proj(i, Hb)
                ops.cos_attn(ws["qkv"], ws["o"], N, T, H, hd)
                ops.gemm_bf16(ws["o"], W.wo[i], X, epilogue=_lib.EPI_RESID, resid=X, gate=mod(i, "gate_a"), ldmod=ld, tokens=T)
                modulate_block(i, "m", X, Hb)
                ops.gemm_bf16(Hb, W.w1[i], ws["u"], epilogue=_lib.EPI_MPSILU)
                ops.gemm_bf16(ws["u"], W.w2[i], X, epilogue=_lib.EPI_RESID, resid=X, gate=mod(i, "gate_m"), ldmod=ld, tokens=T)
                modulate_next(i, X, Hb)
            else:
                ops.gemm_f32(Hb, W.wqkv[i], out=ws["qkv"])
                if cosine:
                    ops.qk_normalize(ws["qkv"], D, hd)
                ops.cos_attn(ws["qkv"], ws["o"], N, T, H, hd)
                ops.gemm_f32(ws["o"], W.wo[i], out=ws["tmp"])
                ops.resid(X, ws["tmp"], X, mod(i, "gate_a"), ld, T)
                modulate_block(i, "m", X, Hb)
                ops.gemm_f32(Hb, W.w1[i], out=ws["u"])
                ops.mp_silu(ws["u"], ws["u"])
                ops.gemm_f32(ws["u"], W.w2[i], out=ws["tmp"])
                ops.resid(X, ws["tmp"], X, mod(i, "gate_m"), ld, T)
                modulate_next(i, X, Hb)
        # ---- final layer (src/blocks/final_layer.py:53-59, src/dit.py:95-100)
        if bf:
            ops.gemm_bf16(Hb, W.wfl, ws["lin"])
        else:
            ops.gemm_f32(Hb, W.wfl, out=ws["lin"])
        out = torch.empty(N, 2 * m.in_channels, m.input_size, m.input_size, device=dev, dtype=torch.float32)
        ops.final_unpatchify(ws["lin"], cb["smu"], cb["ssg"], out, m.patch_size)
        return out
