"""B200-native ``UNetSpatioTemporalConditionModel`` (host-side mirror of the reference interface).

Same constructor data (a reference ``state_dict``), same ``forward`` signature and result layout as
``UNetSpatioTemporalConditionModel.forward`` (animation/modules/unet.py:231-362); every arithmetic op
runs in ``libsa_b200.so`` through :mod:`stableanimator_b200.ops` -- PyTorch only owns device memory
and streams.  There is no fallback: a missing library raises.

Layout: one canonical channel-innermost bf16 activation layout ``[B*F, h, w, C]`` (== tokens
``[B*F*S, C]``).  Spatial ops see frames as the batch, temporal ops index the same buffer with the
frame stride -- the reference's NCHW<->NHWC transposes (unet.py:661,702) and
``[B*F,S,C]<->[B*S,F,C]`` permutes (:931-933,:954-956) never materialise.

Unlike the reference (B == 1 only, unet.py:455) any batch ``B`` is accepted; batch elements are
independent (used for the fused CFG cond/uncond pair, with the pose residual applied to the batch
elements ``>= pose_batch_from`` only, cf. inference_pipeline_animation.py:663-684).
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .arch import ResSpec, TstSpec, UNetConfig, build_stages, unet_param_shapes


def _bf16(t: torch.Tensor, dev, dtype=torch.bfloat16) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=dtype).contiguous()


def _f32(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


class _Ctx:
    """Per-call context-derived tensors (they depend on encoder_hidden_states only)."""
    __slots__ = ("all", "wv", "tconst", "b", "zero_prefix", "ffaddv")


class _StatsArena:
    """fp64 [n, 32, 2] GroupNorm-statistics buffers of ONE UNet call, carved from chunks that are zeroed with one fill each
    (``gemm_tap(gn_stats_zeroed=self.stats_arena)``): inside a CUDA graph that is one memset node for the whole call instead of one
    in front of every convolution -- a memset node between two kernels also keeps the convolution from starting
    under its predecessor's tail (SA_PDL)."""
    CHUNK = 1 << 18  # doubles (2 MB): a B=2, F=16 call uses ~0.8 MB

    def __init__(self, dev):
        self.dev, self.buf, self.used = dev, None, 0

    def take(self, n: int) -> torch.Tensor:
        need = n * 64
        if self.buf is None or self.used + need > self.buf.numel():
            self.buf = torch.zeros(max(self.CHUNK, need), device=self.dev, dtype=torch.float64)
            self.used = 0
        out = self.buf[self.used:self.used + need].view(n, 32, 2)
        self.used += need
        return out


class UNetB200:
    """Weights packed for the kernels + the forward orchestration."""

    def __init__(self, cfg: UNetConfig, state_dict: Dict[str, torch.Tensor], device="cuda", dtype=torch.bfloat16):
        """``dtype``: the 16-bit format of activations and tensor-core weights -- ``torch.bfloat16`` (BASELINE's GPU
        configs) or ``torch.float16``, the reference's default inference dtype (inference_basic.py:486-503).  The
        skinny fp32-activation kernels (time / context embeddings) keep bf16 weights in both modes."""
        if dtype not in ops.ACT_DTYPES:
            raise ValueError(f"dtype must be torch.bfloat16 or torch.float16, got {dtype}")
        self.cfg = cfg
        self.dtype = dtype
        self.dev = torch.device(device)
        self.stages = build_stages(cfg)
        shapes = unet_param_shapes(cfg)
        missing = [k for k in shapes if k not in state_dict]
        extra = [k for k in state_dict if k not in shapes]
        if missing or extra:  # strict=True semantics (inference_basic.py:465-468)
            raise KeyError(f"state_dict mismatch: missing {missing[:5]} unexpected {extra[:5]}")
        for k, shp in shapes.items():
            if tuple(state_dict[k].shape) != tuple(shp):
                raise ValueError(f"{k}: shape {tuple(state_dict[k].shape)} != {shp}")
        self.w: Dict[str, torch.Tensor] = {}
        # GroupNorm statistics accumulated by the producing convolution's epilogue instead of a separate statistics
        # pass (82 of the 105 GroupNorm sites then run the normalise pass only).  B-invariant (tests/test_gemm_gpu.py).
        # Round 1's version was 0.5 ms slower than the separate pass (~400 k same-address fp64 atomics per launch); the
        # round-2 epilogue (packed fp32 column sums, one fp64 atomic set per tile and group) is 0.45 ms faster per B=2
        # call at 64x64x16 (tools/unet_env_ab.py fuse_gn_stats).  SA_GN_FUSE=0: two-pass statistics, for A/B
        self.fuse_gn_stats = os.environ.get("SA_GN_FUSE", "1") == "1"
        # one fill of all the statistics buffers per call instead of a memset in front of every convolution (A/B: False)
        self.stats_arena = os.environ.get("SA_STATS_ARENA", "1") == "1"
        # zero-context (unconditional CFG branch) shortcut of the ID cross-attention; SA_SKIP_ZERO_CTX=0 for A/B
        self.skip_zero_context = os.environ.get("SA_SKIP_ZERO_CTX", "1") != "0"
        # LayerNorm + GEGLU feed-forward + residual (+ blend) as ONE kernel where the channel count allows it (C <= 320: the
        # 64x64 / 128x72 level, where the unfused chain is bound by the GEGLU epilogue and the [M, 4C] hidden tensor's HBM
        # round trip); SA_FF_FUSED=0: LayerNorm -> GEGLU GEMM -> out-projection GEMM, for A/B
        self.fuse_ff = os.environ.get("SA_FF_FUSED", "1") != "0"
        # stride-2 Downsample2D through TMA element strides (SA_STRIDED_DOWN=0: im2col_s2 + GEMM, for A/B)
        self.strided_downsample = os.environ.get("SA_STRIDED_DOWN", "1") != "0"
        # Upsample2D as four phase convolutions (4/9 of the MACs; SA_UP_PHASES=0: upsample2x + 3x3 conv, for A/B)
        self.upsample_phases = os.environ.get("SA_UP_PHASES", "1") != "0"
        self._pos_cache: Dict[Tuple[str, int, int], torch.Tensor] = {}
        self._pack(state_dict)

    # ------------------------------------------------------------------ weight packing
    def _pack(self, sd: Dict[str, torch.Tensor]) -> None:
        dev, w, cfg = self.dev, self.w, self.cfg
        act = self.dtype   # format of the tensor-core weights; the skinny fp32-activation kernels always take bf16

        def lin(name, bias=True, small=False):
            w[f"{name}.w"] = _bf16(sd[f"{name}.weight"], dev, torch.bfloat16 if small else act)
            if bias:
                w[f"{name}.b"] = _f32(sd[f"{name}.bias"], dev)

        def norm(name):
            w[f"{name}.g"] = _f32(sd[f"{name}.weight"], dev)
            w[f"{name}.b"] = _f32(sd[f"{name}.bias"], dev)

        def conv3(name):  # OIHW -> [O, (ky, kx, ci)]
            t = sd[f"{name}.weight"]
            w[f"{name}.w"] = _bf16(t.permute(0, 2, 3, 1).reshape(t.shape[0], -1), dev, act)
            w[f"{name}.b"] = _f32(sd[f"{name}.bias"], dev)

        def convt(name):  # OIDHW (3,1,1) -> [O, (kt, ci)]
            t = sd[f"{name}.weight"][:, :, :, 0, 0]
            w[f"{name}.w"] = _bf16(t.permute(0, 2, 1).reshape(t.shape[0], -1), dev, act)
            w[f"{name}.b"] = _f32(sd[f"{name}.bias"], dev)

        def geglu(name):
            wi, bi = ops.interleave_geglu(sd[f"{name}.net.0.proj.weight"], sd[f"{name}.net.0.proj.bias"])
            w[f"{name}.w1"], w[f"{name}.b1"] = _bf16(wi, dev, act), _f32(bi, dev)
            w[f"{name}.w2"] = _bf16(sd[f"{name}.net.2.weight"], dev, act)
            w[f"{name}.b2"] = _f32(sd[f"{name}.net.2.bias"], dev)

        def qkv(name):
            cat = torch.cat([sd[f"{name}.to_q.weight"], sd[f"{name}.to_k.weight"], sd[f"{name}.to_v.weight"]], 0)
            w[f"{name}.qkv"] = _bf16(cat, dev, act)
            lin(f"{name}.to_out.0")

        # conv_in: direct conv weights [k*k*cin, cout]
        t = sd["conv_in.weight"]
        w["conv_in.w"] = _bf16(t.permute(2, 3, 1, 0).reshape(-1, t.shape[0]), dev, act)
        w["conv_in.b"] = _f32(sd["conv_in.bias"], dev)
        if t.shape[1] == 8 and t.shape[2:] == (3, 3):  # tensor-core path: [O, 72 (ky, kx, ci) | 56 zeros]
            wg = torch.zeros(t.shape[0], 128, dtype=torch.float32)
            wg[:, :72] = t.permute(0, 2, 3, 1).reshape(t.shape[0], 72).float().cpu()
            w["conv_in.wg"] = _bf16(wg, dev, act)
        for n in ("time_embedding.linear_1", "time_embedding.linear_2", "add_embedding.linear_1",
                  "add_embedding.linear_2"):
            lin(n, small=True)

        temb_w, temb_b, self.temb_off = [], [], {}
        ctx_w, self.ctx_off = [], {}
        off_t = off_c = 0
        for st in self.stages:
            for r in st.res:
                s_, t_ = f"{r.name}.spatial_res_block", f"{r.name}.temporal_res_block"
                norm(f"{s_}.norm1"); conv3(f"{s_}.conv1"); norm(f"{s_}.norm2"); conv3(f"{s_}.conv2")
                if r.shortcut:
                    t = sd[f"{s_}.conv_shortcut.weight"]
                    w[f"{s_}.conv_shortcut.w"] = _bf16(t.reshape(t.shape[0], t.shape[1]), dev, act)
                    w[f"{s_}.conv_shortcut.b"] = _f32(sd[f"{s_}.conv_shortcut.bias"], dev)
                norm(f"{t_}.norm1"); convt(f"{t_}.conv1"); norm(f"{t_}.norm2"); convt(f"{t_}.conv2")
                for blk in (s_, t_):  # all 44 time_emb_proj share the input silu(emb): one skinny GEMM
                    temb_w.append(sd[f"{blk}.time_emb_proj.weight"])
                    temb_b.append(sd[f"{blk}.time_emb_proj.bias"])
                    self.temb_off[blk] = off_t
                    off_t += r.cout
                w[f"{r.name}.alpha"] = float(torch.sigmoid(sd[f"{r.name}.time_mixer.mix_factor"].float()).item())
            for t in st.tst:
                if t is None:
                    continue
                c = t.c
                norm(f"{t.name}.norm"); lin(f"{t.name}.proj_in"); lin(f"{t.name}.proj_out")
                b = f"{t.name}.transformer_blocks.0"
                norm(f"{b}.norm1"); qkv(f"{b}.attn1"); norm(f"{b}.norm2")
                lin(f"{b}.attn2.to_q", bias=False); lin(f"{b}.attn2.to_out.0")
                # the same matrix for sa_small_linear (W_out . v_clip): bf16 in both modes
                w[f"{b}.attn2.to_out.0.w_sl"] = (w[f"{b}.attn2.to_out.0.w"] if act == torch.bfloat16 else
                                                 _bf16(sd[f"{b}.attn2.to_out.0.weight"], dev))
                # rowsum(W_out) for the folded distribution alignment (weight-only constant)
                w[f"{b}.attn2.rowsum"] = w[f"{b}.attn2.to_out.0.w"].float().sum(dim=1).contiguous()
                norm(f"{b}.norm3"); geglu(f"{b}.ff")
                tb = f"{t.name}.temporal_transformer_blocks.0"
                norm(f"{tb}.norm_in"); geglu(f"{tb}.ff_in"); norm(f"{tb}.norm1"); qkv(f"{tb}.attn1")
                lin(f"{tb}.attn2.to_out.0", small=True); norm(f"{tb}.norm3"); geglu(f"{tb}.ff")
                lin(f"{t.name}.time_pos_embed.linear_1", small=True); lin(f"{t.name}.time_pos_embed.linear_2", small=True)
                w[f"{t.name}.alpha"] = float(torch.sigmoid(sd[f"{t.name}.time_mixer.mix_factor"].float()).item())
                # context projections of the 5 tokens, batched over the whole net:
                #   [spatial to_v | id_to_k | id_to_v | temporal to_v]  (to_k of a single key is inert)
                ctx_w += [sd[f"{b}.attn2.to_v.weight"], sd[f"{b}.attn2.processor.id_to_k.weight"],
                          sd[f"{b}.attn2.processor.id_to_v.weight"], sd[f"{tb}.attn2.to_v.weight"]]
                self.ctx_off[t.name] = off_c
                off_c += 4 * c
            for key in (st.downsample, st.upsample):
                if key:
                    conv3(key)
            if st.upsample:  # Upsample2D as four 2x2-tap phase convolutions on the low-resolution input
                for i, (_, _, _, wp) in enumerate(ops.upsample_conv_phases(sd[f"{st.upsample}.weight"])):
                    w[f"{st.upsample}.ph{i}.w"] = _bf16(wp, dev, act)
        w["temb_all.w"] = _bf16(torch.cat(temb_w, 0), dev)
        w["temb_all.b"] = _f32(torch.cat(temb_b, 0), dev)
        w["ctx_all.w"] = _bf16(torch.cat(ctx_w, 0), dev)
        self.n_temb, self.n_ctx = off_t, off_c
        norm("conv_norm_out")
        t = sd["conv_out.weight"]  # pad N 4 -> 8 (zero rows) for the 8-column store granularity
        wo = torch.zeros(8, t.shape[1] * 9, dtype=torch.float32)
        wo[: t.shape[0]] = t.permute(0, 2, 3, 1).reshape(t.shape[0], -1).float().cpu()
        bo = torch.zeros(8)
        bo[: t.shape[0]] = sd["conv_out.bias"].float().cpu()
        w["conv_out.w"], w["conv_out.b"] = _bf16(wo, dev, act), _f32(bo, dev)

    # ------------------------------------------------------------------ embeddings / context
    def _time_embedding(self, timestep: torch.Tensor, added_time_ids: torch.Tensor, B: int) -> torch.Tensor:
        """unet.py:249-271 -> fp32 [B, 1280]; then all time_emb_proj(silu(emb)) at once -> [B, n_temb]."""
        w, cfg = self.w, self.cfg
        t = timestep.reshape(-1).to(self.dev, torch.float32)
        t = t.expand(B).contiguous()
        e1 = ops.small_linear(ops.sinusoid(t, cfg.block_out_channels[0]), w["time_embedding.linear_1.w"],
                              w["time_embedding.linear_1.b"], silu_out=True)
        emb = ops.small_linear(e1, w["time_embedding.linear_2.w"], w["time_embedding.linear_2.b"])
        ids = added_time_ids.to(self.dev, torch.float32).reshape(-1).contiguous()
        te = ops.sinusoid(ids, cfg.addition_time_embed_dim).reshape(B, -1)
        a1 = ops.small_linear(te, w["add_embedding.linear_1.w"], w["add_embedding.linear_1.b"], silu_out=True)
        ops.small_linear(a1, w["add_embedding.linear_2.w"], w["add_embedding.linear_2.b"], accumulate=True, out=emb)
        return ops.small_linear(emb, w["temb_all.w"], w["temb_all.b"], silu_in=True)

    def prepare_context(self, encoder_hidden_states: torch.Tensor) -> _Ctx:
        """Everything that depends on the conditioning tokens only (constant over the 25 steps and all
        frame tiles): K/V projections of the CLIP + 4 face tokens for every attention layer, W_out.v_clip
        for the folded alignment, and the temporal attn2 output (softmax over ONE key == V)."""
        w = self.w
        ehs = encoder_hidden_states.to(self.dev, torch.float32).contiguous()
        B, L, D = ehs.shape
        if L != 1 + self.cfg.num_tokens or D != self.cfg.cross_attention_dim:
            raise ValueError(f"encoder_hidden_states must be [B, {1 + self.cfg.num_tokens}, "
                             f"{self.cfg.cross_attention_dim}], got {tuple(ehs.shape)}")
        ctx = _Ctx()
        ctx.b = B
        # leading batch elements whose conditioning is exactly zero (the unconditional CFG branch,
        # inference_pipeline_animation.py:232-241, :526-533): for them the ID cross-attention is the constant to_out bias
        # (k = v = 0 -> both attention branches and all four alignment statistics are 0, unet.py:1999-2024), so
        # _tst skips norm2 / to_q / attention / to_out on their rows.  One host read per video, none per step.
        nz = (ehs.abs().amax(dim=(1, 2)) > 0).tolist()
        ctx.zero_prefix = next((i for i, v in enumerate(nz) if v), B) if self.skip_zero_context else 0
        ctx.all = ops.small_linear(ehs.reshape(B * L, D), w["ctx_all.w"]).reshape(B, L, self.n_ctx)
        ctx.wv, ctx.tconst, ctx.ffaddv = {}, {}, {}
        for st in self.stages:
            for t in st.tst:
                if t is None:
                    continue
                o, c = self.ctx_off[t.name], t.c
                b = f"{t.name}.transformer_blocks.0"
                tb = f"{t.name}.temporal_transformer_blocks.0"
                vclip = ctx.all[:, 0, o:o + c]
                tv = ctx.all[:, 0, o + 3 * c:o + 4 * c]
                ctx.wv[t.name] = ops.small_linear(vclip, w[f"{b}.attn2.to_out.0.w_sl"])
                ctx.tconst[t.name] = ops.small_linear(tv, w[f"{tb}.attn2.to_out.0.w"], w[f"{tb}.attn2.to_out.0.b"])
                if ctx.zero_prefix > 0:   # attn2 output of the zero-context batch elements: the to_out bias (others: 0)
                    av = torch.zeros((B, c), device=self.dev, dtype=torch.float32)
                    av[:ctx.zero_prefix] = w[f"{b}.attn2.to_out.0.b"]
                    ctx.ffaddv[t.name] = av
        return ctx

    def _pos_embed(self, t: TstSpec, B: int, F: int) -> torch.Tensor:
        """time_pos_embed(sinusoid(arange(F))) (unet.py:666-677): weight-only constant, fp32 [B*F, C]."""
        key = (t.name, B, F)
        if key not in self._pos_cache:
            w = self.w
            pos = torch.arange(F, device=self.dev, dtype=torch.float32).repeat(B).contiguous()
            h = ops.small_linear(ops.sinusoid(pos, t.c), w[f"{t.name}.time_pos_embed.linear_1.w"],
                                 w[f"{t.name}.time_pos_embed.linear_1.b"], silu_out=True)
            self._pos_cache[key] = ops.small_linear(h, w[f"{t.name}.time_pos_embed.linear_2.w"],
                                                    w[f"{t.name}.time_pos_embed.linear_2.b"])
        return self._pos_cache[key]

    # ------------------------------------------------------------------ blocks
    def _res_block(self, r: ResSpec, x0, x1, temb, B, F, h, wd, arena, x0_stats=None, out_stats=None):
        """SpatioTemporalResBlock.forward (unet.py:1579-1608) on NHWC tokens.  x0 [M, C0] (+ skip x1).
        The statistics of the three inner GroupNorms are accumulated by the epilogue of the convolution that
        produces their input (sa_gemm_args.gn_stats), so each of them is a single normalise pass.  ``x0_stats``:
        per-frame statistics of ``x0`` from its producer (single-source input only); ``out_stats``: fp64
        [B*F, 32, 2] to fill with the per-frame statistics of the block's output for the GroupNorm that follows."""
        w = self.w
        S = h * wd
        s_, t_ = f"{r.name}.spatial_res_block", f"{r.name}.temporal_res_block"
        grid2d, gridt = (wd, h, B * F), (S, F, B)
        fuse = self.fuse_gn_stats
        st_f = arena.take(B * F) if fuse else None   # per frame
        st_t1 = arena.take(B) if fuse else None      # whole tile
        st_t2 = arena.take(B) if fuse else None
        n1 = ops.groupnorm(x0, w[f"{s_}.norm1.g"], w[f"{s_}.norm1.b"], r.eps, S, x1=x1, silu=True,
                           stats_ws=x0_stats, stats_ready=x0_stats is not None)
        o_s = self.temb_off[s_]
        h1 = ops.gemm_tap(n1, w[f"{s_}.conv1.w"], grid=grid2d, taps=ops.TAPS_3X3, bias=w[f"{s_}.conv1.b"],
                          row_bias=temb[:, o_s:o_s + r.cout], rows_per_group=F * S, gn_stats=st_f, gn_rows_per_stat=S,
                          gn_stats_zeroed=self.stats_arena)
        n2 = ops.groupnorm(h1, w[f"{s_}.norm2.g"], w[f"{s_}.norm2.b"], r.eps, S, silu=True, stats_ws=st_f,
                           stats_ready=fuse)
        if r.shortcut:
            sc = ops.gemm_tap(x0, w[f"{s_}.conv_shortcut.w"], a1=x1, bias=w[f"{s_}.conv_shortcut.b"])
        else:
            sc = x0
        xs = ops.gemm_tap(n2, w[f"{s_}.conv2.w"], grid=grid2d, taps=ops.TAPS_3X3, bias=w[f"{s_}.conv2.b"],
                          residual=sc, gn_stats=st_t1, gn_rows_per_stat=F * S, gn_stats_zeroed=self.stats_arena)
        # temporal: GroupNorm statistics over the whole tile (5-D input, unet.py:1749), 3-tap conv over frames
        n3 = ops.groupnorm(xs, w[f"{t_}.norm1.g"], w[f"{t_}.norm1.b"], r.eps, F * S, silu=True, stats_ws=st_t1,
                           stats_ready=fuse)
        o_t = self.temb_off[t_]
        h2 = ops.gemm_tap(n3, w[f"{t_}.conv1.w"], grid=gridt, taps=ops.TAPS_T3, bias=w[f"{t_}.conv1.b"],
                          row_bias=temb[:, o_t:o_t + r.cout], rows_per_group=F * S, gn_stats=st_t2,
                          gn_rows_per_stat=F * S, gn_stats_zeroed=self.stats_arena)
        n4 = ops.groupnorm(h2, w[f"{t_}.norm2.g"], w[f"{t_}.norm2.b"], r.eps, F * S, silu=True, stats_ws=st_t2,
                           stats_ready=fuse)
        # conv2 + residual (xs), then AlphaBlender: alpha*xs + (1-alpha)*(xs + conv2) == xs + (1-alpha)*conv2,
        # one epilogue with a single residual read
        return ops.gemm_tap(n4, w[f"{t_}.conv2.w"], grid=gridt, taps=ops.TAPS_T3, bias=w[f"{t_}.conv2.b"],
                            residual=xs, out_scale=1.0 - w[f"{r.name}.alpha"], gn_stats=out_stats, gn_rows_per_stat=S,
                            gn_stats_zeroed=self.stats_arena)

    def _tst(self, t: TstSpec, x, ctx: _Ctx, B, F, h, wd, x_stats=None):
        """TransformerSpatioTemporalModel.forward (unet.py:623-708) on NHWC tokens x [M, C].  ``x_stats``: per-frame
        GroupNorm statistics of ``x`` accumulated by the res block that produced it."""
        w = self.w
        S, c, heads = h * wd, t.c, t.heads
        BF = B * F
        b = f"{t.name}.transformer_blocks.0"
        tb = f"{t.name}.temporal_transformer_blocks.0"
        n = ops.groupnorm(x, w[f"{t.name}.norm.g"], w[f"{t.name}.norm.b"], 1e-6, S, silu=False, stats_ws=x_stats,
                          stats_ready=x_stats is not None)
        y = ops.gemm_tap(n, w[f"{t.name}.proj_in.w"], bias=w[f"{t.name}.proj_in.b"])
        # --- BasicTransformerBlock (unet.py:525-556)
        a = ops.layernorm(y, w[f"{b}.norm1.g"], w[f"{b}.norm1.b"])
        qkv = ops.gemm_tap(a, w[f"{b}.attn1.qkv"], f16_from_col=2 * c)  # v block in fp16 for the P V MMA
        o = ops.attn_spatial(qkv, BF, S, heads)
        y1 = ops.gemm_tap(o, w[f"{b}.attn1.to_out.0.w"], bias=w[f"{b}.attn1.to_out.0.b"], residual=y)
        # attn2 (ID cross-attention) on the batch elements with a non-zero context; for the zero-context prefix
        # (rows [0, mz)) the block adds the constant to_out bias, fused into norm3 below
        zb = ctx.zero_prefix
        mz, bn = zb * F * S, B - zb
        off = self.ctx_off[t.name]
        nt = self.cfg.num_tokens
        ff_fused = self.fuse_ff and c in ops.FF_FUSED_CHANNELS
        # fused path: the attn2 output is added to y1 IN PLACE on the rows that have a context (each epilogue thread reads
        # the residual element it overwrites), so one feed-forward launch covers all rows
        y2 = y1 if ff_fused else torch.empty_like(y1)
        a = None if ff_fused else torch.empty_like(y1)
        if bn > 0:
            an = ops.layernorm(y1[mz:], w[f"{b}.norm2.g"], w[f"{b}.norm2.b"])
            q = ops.gemm_tap(an, w[f"{b}.attn2.to_q.w"])
            ik = ctx.all[zb:, 1:1 + nt, off + c:off + 2 * c]
            iv = ctx.all[zb:, 1:1 + nt, off + 2 * c:off + 3 * c]
            stats = torch.empty((bn * F, 2), device=self.dev, dtype=torch.float64)
            ip = ops.attn_id(q, ik, iv, bn, F, S, heads, stats)
            rs = torch.empty((bn * F,), device=self.dev, dtype=torch.float32)
            rb = torch.empty((bn * F, c), device=self.dev, dtype=torch.float32)
            ops.attn_id_finalize(stats, ctx.all[zb:, 0, off:off + c], ctx.wv[t.name][zb:], w[f"{b}.attn2.rowsum"],
                                 w[f"{b}.attn2.to_out.0.b"], bn, F, S, c, rs, rb)
            ops.gemm_tap(ip, w[f"{b}.attn2.to_out.0.w"], row_scale=rs, row_bias=rb, rows_per_group=S,
                         residual=y1[mz:], out=y2[mz:])
            if not ff_fused:
                ops.layernorm(y2[mz:], w[f"{b}.norm3.g"], w[f"{b}.norm3.b"], out=a[mz:])
        if ff_fused:
            # norm3 -> ff -> + residual in one kernel; on the zero-context batch elements the attn2 output (the constant
            # to_out bias) enters as the per-batch-element additive vector, zero on the others
            y3 = ops.ff_fused(y2, w[f"{b}.norm3.g"], w[f"{b}.norm3.b"], w[f"{b}.ff.w1"], w[f"{b}.ff.b1"], w[f"{b}.ff.w2"],
                              w[f"{b}.ff.b2"], addv=ctx.ffaddv.get(t.name), rows_per_group=F * S)
        else:
            if zb > 0:  # y2 = y1 + to_out.bias, a = norm3(y2) in one pass
                ops.layernorm(y1[:mz], w[f"{b}.norm3.g"], w[f"{b}.norm3.b"], addv=w[f"{b}.attn2.to_out.0.b"].reshape(1, c),
                              rows_per_group=mz, x_out=y2[:mz], out=a[:mz])
            g = ops.gemm_tap(a, w[f"{b}.ff.w1"], bias=w[f"{b}.ff.b1"], geglu=True)
            y3 = ops.gemm_tap(g, w[f"{b}.ff.w2"], bias=w[f"{b}.ff.b2"], residual=y2)
        # --- TemporalBasicTransformerBlock (unet.py:919-958) on hidden + frame-position embedding (:686)
        if ff_fused:
            z1 = ops.ff_fused(y3, w[f"{tb}.norm_in.g"], w[f"{tb}.norm_in.b"], w[f"{tb}.ff_in.w1"], w[f"{tb}.ff_in.b1"],
                              w[f"{tb}.ff_in.w2"], w[f"{tb}.ff_in.b2"], addv=self._pos_embed(t, B, F), rows_per_group=S)
        else:
            z0 = torch.empty_like(y3)
            a = ops.layernorm(y3, w[f"{tb}.norm_in.g"], w[f"{tb}.norm_in.b"], addv=self._pos_embed(t, B, F),
                              rows_per_group=S, x_out=z0)
            g = ops.gemm_tap(a, w[f"{tb}.ff_in.w1"], bias=w[f"{tb}.ff_in.b1"], geglu=True)
            z1 = ops.gemm_tap(g, w[f"{tb}.ff_in.w2"], bias=w[f"{tb}.ff_in.b2"], residual=z0)
        a = ops.layernorm(z1, w[f"{tb}.norm1.g"], w[f"{tb}.norm1.b"])
        qkv = ops.gemm_tap(a, w[f"{tb}.attn1.qkv"])
        o = ops.attn_temporal(qkv, B, F, S, heads)
        z2 = ops.gemm_tap(o, w[f"{tb}.attn1.to_out.0.w"], bias=w[f"{tb}.attn1.to_out.0.b"], residual=z1)
        # attn2 vs the single first-frame CLIP token == constant vector per batch element (:945-947)
        if ff_fused:
            z4 = ops.ff_fused(z2, w[f"{tb}.norm3.g"], w[f"{tb}.norm3.b"], w[f"{tb}.ff.w1"], w[f"{tb}.ff.b1"], w[f"{tb}.ff.w2"],
                              w[f"{tb}.ff.b2"], addv=ctx.tconst[t.name], rows_per_group=F * S, blend_x=y3,
                              alpha=w[f"{t.name}.alpha"])
        else:
            z3 = torch.empty_like(z2)
            a = ops.layernorm(z2, w[f"{tb}.norm3.g"], w[f"{tb}.norm3.b"], addv=ctx.tconst[t.name],
                              rows_per_group=F * S, x_out=z3)
            g = ops.gemm_tap(a, w[f"{tb}.ff.w1"], bias=w[f"{tb}.ff.b1"], geglu=True)
            # ff out + residual, then time_mixer: alpha * spatial(y3) + (1 - alpha) * temporal  (:693-697)
            z4 = ops.gemm_tap(g, w[f"{tb}.ff.w2"], bias=w[f"{tb}.ff.b2"], residual=z3, blend_x=y3,
                              alpha=w[f"{t.name}.alpha"])
        return ops.gemm_tap(z4, w[f"{t.name}.proj_out.w"], bias=w[f"{t.name}.proj_out.b"], residual=x)

    def _conv_in(self, x_in: torch.Tensor, pose: Optional[torch.Tensor], pose_from: int) -> torch.Tensor:
        """conv_in (unet.py:283) + pose residual on the frames >= pose_from (:284-286) -> tokens [M, C0].
        K = 72 is gathered to a [M, 128] operand and runs on the tensor-core GEMM (two launches when only
        part of the batch carries the pose residual); other conv_in geometries use the direct kernel."""
        w, c0 = self.w, self.cfg.block_out_channels[0]
        n, h, wd, cin = x_in.shape
        if "conv_in.wg" not in w or c0 % 64:
            return ops.conv_direct(x_in, w["conv_in.w"], w["conv_in.b"], c0, 3, 1, 1, add=pose,
                                   add_from=pose_from).reshape(-1, c0)
        cols = ops.im2col3x3_c8(x_in.contiguous())
        x = torch.empty((n * h * wd, c0), device=self.dev, dtype=self.dtype)
        r0 = pose_from * h * wd if pose is not None else n * h * wd
        if r0 > 0:
            ops.gemm_tap(cols[:r0], w["conv_in.wg"], bias=w["conv_in.b"], out=x[:r0])
        if pose is not None:
            ops.gemm_tap(cols[r0:], w["conv_in.wg"], bias=w["conv_in.b"], residual=pose.reshape(-1, c0), out=x[r0:])
        return x

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward_nhwc(self, x_in: torch.Tensor, timestep: torch.Tensor, ctx: _Ctx, added_time_ids: torch.Tensor,
                     B: int, F: int, h: int, wd: int, pose_nhwc: Optional[torch.Tensor] = None,
                     pose_batch_from: int = 0, taps: Optional[dict] = None) -> torch.Tensor:
        """x_in bf16 [B*F, h, w, 8] -> fp32 [B*F*h*w, 8] (first 4 columns = prediction).
        ``pose_nhwc`` bf16 [(B - pose_batch_from)*F, h, w, C0] is added after conv_in to the batch
        elements >= pose_batch_from (unet.py:284-286)."""
        w, cfg = self.w, self.cfg
        if h % 8 or wd % 8:
            raise ValueError("latent height/width must be multiples of 8 (three stride-2 stages)")
        if ctx.b != B:
            raise ValueError("context batch does not match")
        temb = self._time_embedding(timestep, added_time_ids, B)
        arena = _StatsArena(self.dev)
        fstats = (lambda: arena.take(B * F)) if self.fuse_gn_stats else (lambda: None)  # per-frame statistics of a block output
        x = self._conv_in(x_in.reshape(B * F, h, wd, cfg.in_channels), pose_nhwc, pose_batch_from * F)
        if taps is not None:
            taps["conv_in"] = x
        skips: List[Tuple[torch.Tensor, int]] = [x]
        ch, cw = h, wd
        for st in self.stages:
            if st.kind == "down":
                for r, t in zip(st.res, st.tst):
                    xst = fstats() if t is not None else None
                    x = self._res_block(r, x, None, temb, B, F, ch, cw, arena, out_stats=xst)
                    if taps is not None:
                        taps[r.name] = x
                    if t is not None:
                        x = self._tst(t, x, ctx, B, F, ch, cw, x_stats=xst)
                        if taps is not None:
                            taps[t.name] = x
                    skips.append(x)
                if st.downsample and self.strided_downsample:
                    # Downsample2D (unet.py:1459-1483): the stride-2 taps come straight from the full-resolution
                    # tensor through TMA element strides -- no im2col buffer
                    x = ops.gemm_tap(x.reshape(B * F, ch, cw, st.cout), w[f"{st.downsample}.w"],
                                     grid=(cw // 2, ch // 2, B * F), taps=ops.TAPS_3X3, in_stride=2,
                                     bias=w[f"{st.downsample}.b"])
                    ch, cw = ch // 2, cw // 2
                    skips.append(x)
                elif st.downsample:
                    cols = ops.im2col_s2(x.reshape(B * F, ch, cw, st.cout))
                    x = ops.gemm_tap(cols, w[f"{st.downsample}.w"], bias=w[f"{st.downsample}.b"])
                    ch, cw = ch // 2, cw // 2
                    skips.append(x)
            elif st.kind == "mid":
                xst = fstats()
                x = self._res_block(st.res[0], x, None, temb, B, F, ch, cw, arena, out_stats=xst)
                x = self._tst(st.tst[0], x, ctx, B, F, ch, cw, x_stats=xst)
                x = self._res_block(st.res[1], x, None, temb, B, F, ch, cw, arena)
                if taps is not None:
                    taps["mid_block"] = x
            else:
                for r, t in zip(st.res, st.tst):
                    xst = fstats() if t is not None else None
                    x = self._res_block(r, x, skips.pop(), temb, B, F, ch, cw, arena, out_stats=xst)
                    if t is not None:
                        x = self._tst(t, x, ctx, B, F, ch, cw, x_stats=xst)
                    if taps is not None:
                        taps[r.name] = x
                if st.upsample and self.upsample_phases:
                    # nearest x2 + conv3x3 (unet.py:1789-1807): each output phase (py, px) is a 2x2-tap convolution of
                    # the low-resolution tensor written to every other pixel -- 4/9 of the MACs, no upsampled tensor
                    xin = x.reshape(B * F, ch, cw, st.cout)
                    up = torch.empty((B * F, 2 * ch, 2 * cw, st.cout), device=self.dev, dtype=self.dtype)
                    for i, (py, px, ptaps) in enumerate(ops.UPSAMPLE_PHASE_TAPS):
                        ops.gemm_tap(xin, w[f"{st.upsample}.ph{i}.w"], grid=(cw, ch, B * F), taps=ptaps,
                                     bias=w[f"{st.upsample}.b"], out=up[:, py::2, px::2, :])
                    ch, cw = ch * 2, cw * 2
                    x = up.reshape(-1, st.cout)
                elif st.upsample:
                    up = ops.upsample2x(x.reshape(B * F, ch, cw, st.cout))
                    ch, cw = ch * 2, cw * 2
                    x = ops.gemm_tap(up, w[f"{st.upsample}.w"], grid=(cw, ch, B * F), taps=ops.TAPS_3X3,
                                     bias=w[f"{st.upsample}.b"])
        n = ops.groupnorm(x, w["conv_norm_out.g"], w["conv_norm_out.b"], 1e-5, ch * cw, silu=True)
        return ops.gemm_tap(n, w["conv_out.w"], grid=(cw, ch, B * F), taps=ops.TAPS_3X3, bias=w["conv_out.b"],
                            out_f32=True)

    @torch.no_grad()
    def forward(self, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor,
                added_time_ids: torch.Tensor, pose_latents: Optional[torch.Tensor] = None,
                image_only_indicator=False, return_dict: bool = True, pose_batch_from: int = 0,
                taps: Optional[dict] = None):
        """Drop-in for ``UNetSpatioTemporalConditionModel.forward`` (unet.py:231-362): NCHW tensors in,
        ``(sample[B, F, 4, h, w],)`` out, in ``sample``'s dtype.  ``image_only_indicator`` and
        ``return_dict`` are accepted and ignored exactly as the reference does (:288, :362)."""
        if sample.dim() != 5 or sample.shape[2] != self.cfg.in_channels:
            raise ValueError(f"sample must be [B, F, {self.cfg.in_channels}, h, w]")
        B, F, _, h, wd = sample.shape
        s = sample.to(self.dev)
        if s.dtype not in (torch.float32, self.dtype):
            s = s.float()
        x_in = ops.nchw_to_nhwc(s.reshape(B * F, self.cfg.in_channels, h, wd).contiguous(), dtype=self.dtype)
        pose = None
        if pose_latents is not None:
            p = pose_latents.to(self.dev)
            if p.dtype not in (torch.float32, self.dtype):
                p = p.float()
            pose = ops.nchw_to_nhwc(p.contiguous(), dtype=self.dtype)
        if not torch.is_tensor(timestep):
            timestep = torch.tensor(float(timestep))
        ctx = self.prepare_context(encoder_hidden_states)
        out = self.forward_nhwc(x_in, timestep, ctx, added_time_ids, B, F, h, wd, pose, pose_batch_from, taps)
        y = ops.nhwc_to_nchw_f32(out, B * F, self.cfg.out_channels, h * wd)
        y = y.reshape(B, F, self.cfg.out_channels, h, wd)
        return (y.to(sample.dtype) if sample.dtype != torch.float32 else y,)

    __call__ = forward
