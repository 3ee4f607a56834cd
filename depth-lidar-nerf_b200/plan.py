"""Static execution plan of one NeRF network for the tcgen05 MLP kernels.

Given the network shape (run_nerf_helpers.py:78-111) this builds, once:
  * the flat fp32 parameter layout (reference registration order, so optimizer / checkpoint
    ordering is unchanged),
  * the forward and dgrad chain programs (``DlnChainProgram``),
  * the weight-pack jobs (fp32 -> bf16 swizzled stages) for both chains,
  * the wgrad items and the stash slot maps.
Everything here is host-side integer bookkeeping; no arithmetic on tensor data.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

from . import _lib as L


def _ceil_div(a: int, b: int) -> int:
    return (a + b - 1) // b


@dataclass
class NetShape:
    D: int = 8
    W: int = 256
    input_ch: int = 63
    input_ch_views: int = 27
    output_ch: int = 5
    skips: Sequence[int] = (4,)
    use_viewdirs: bool = True
    semantic_num_classes: int = 0      # K > 0: semantic_linear = Linear(W, W/2) -> Linear(W/2, K) (helpers:107-111)

    def validate(self) -> None:
        if self.W % 64 or not 64 <= self.W <= 256:
            raise NotImplementedError("netwidth must be 64, 128, 192 or 256 (got %d): the sm_100a MLP kernels compute 256 "
                                      "output columns per layer, narrower layers are zero-padded up to that" % self.W)
        if self.W != 256 and self.semantic_num_classes:
            raise NotImplementedError("the semantic-head kernels are built for netwidth 256")
        if self.D < 1 or self.D > (8 if self.use_viewdirs else 9):
            raise NotImplementedError("netdepth must be in [1, %d] (bias staging in shared memory)"
                                      % (8 if self.use_viewdirs else 9))
        if self.input_ch > 64 or self.input_ch_views > 64:
            raise NotImplementedError("encoded inputs wider than 64 channels are not supported")
        if (self.input_ch - 3) % 6 or (self.use_viewdirs and (self.input_ch_views - 3) % 6):
            raise NotImplementedError("input widths must be 3+6L (get_embedder)")
        for s in self.skips:
            if s == self.D - 1:
                raise ValueError("a skip after the last layer changes alpha_linear's fan-in; the reference "
                                 "module cannot be built that way either (run_nerf_helpers.py:91,:119)")
        if not self.use_viewdirs and not (1 <= self.output_ch <= 5):
            raise NotImplementedError("output_ch must be in [1,5]")
        if not 0 <= self.semantic_num_classes <= L.SEM_MAX_CLASSES:
            raise NotImplementedError("the semantic head kernels handle up to %d classes" % L.SEM_MAX_CLASSES)

    @property
    def sem_K(self) -> int:
        """Semantic logits per output row: the reference only evaluates the head with view directions
        (run_nerf_helpers.py:122-140); without them the layers exist but are never used."""
        return self.semantic_num_classes if self.use_viewdirs else 0

    @property
    def L_pts(self) -> int:
        return (self.input_ch - 3) // 6

    @property
    def L_dir(self) -> int:
        return (self.input_ch_views - 3) // 6 if self.use_viewdirs else 0

    @property
    def out_ch(self) -> int:
        """Channels of the network output row."""
        return 4 if self.use_viewdirs else self.output_ch

    def fan_in(self, i: int) -> int:
        if i == 0:
            return self.input_ch
        return self.W + self.input_ch if (i - 1) in self.skips else self.W

    def param_shapes(self) -> List[Tuple[str, Tuple[int, ...]]]:
        """Registration order of the reference module (run_nerf_helpers.py:90-111)."""
        out: List[Tuple[str, Tuple[int, ...]]] = []
        for i in range(self.D):
            out.append(("pts_linears.%d.weight" % i, (self.W, self.fan_in(i))))
            out.append(("pts_linears.%d.bias" % i, (self.W,)))
        out.append(("views_linears.0.weight", (self.W // 2, self.input_ch_views + self.W)))
        out.append(("views_linears.0.bias", (self.W // 2,)))
        if self.use_viewdirs:
            out += [("feature_linear.weight", (self.W, self.W)), ("feature_linear.bias", (self.W,)),
                    ("alpha_linear.weight", (1, self.W)), ("alpha_linear.bias", (1,)),
                    ("rgb_linear.weight", (3, self.W // 2)), ("rgb_linear.bias", (3,))]
        else:
            out += [("output_linear.weight", (self.output_ch, self.W)), ("output_linear.bias", (self.output_ch,))]
        if self.semantic_num_classes:
            out += [("semantic_linear.0.weight", (self.W // 2, self.W)), ("semantic_linear.0.bias", (self.W // 2,)),
                    ("semantic_linear.1.weight", (self.semantic_num_classes, self.W // 2)),
                    ("semantic_linear.1.bias", (self.semantic_num_classes,))]
        return out


@dataclass
class Plan:
    shape: NetShape
    offsets: Dict[str, int] = field(default_factory=dict)   # float offsets into the flat parameter buffer
    n_params: int = 0
    fwd: L.ChainProgram = None
    bwd: L.ChainProgram = None
    fwd_jobs: List[L.PackJob] = field(default_factory=list)
    bwd_jobs: List[L.PackJob] = field(default_factory=list)
    fwd_blob_bytes: int = 0
    bwd_blob_bytes: int = 0
    wgrad: List[L.WgradItem] = field(default_factory=list)
    fwd_slots: int = 0
    bwd_slots: int = 0
    mask_slots: int = 0
    # feature_linear folded into views_linears (see build_plan): derived operands live behind the parameters in the
    # flat buffer ([0, n_params) parameters | M [W/2 x W] | b' [W/2]); the flat gradient buffer has the same layout,
    # its tail being per-call scratch (dM, db') that dln_mlp_unfold_grads turns into the real gradients
    fold: bool = False
    n_flat: int = 0
    off_M: int = -1
    off_bM: int = -1
    # semantic head (csrc/semantic_kernels.cu): derived operands A, a, Sw, sc behind M / b' in the flat buffer; the
    # gradient buffer's copies are per-call scratch (dA, da, dSw, dsc) that dln_sem_unfold_grads consumes
    sem: L.SemOffsets = None
    h_last_slot: int = -1


def _set_k(step: L.ChainStep, slabs: List[int], cnts: List[int]) -> None:
    step.nk = len(slabs)
    for j, (s, c) in enumerate(zip(slabs, cnts)):
        step.kslab[j] = s
        step.kcnt[j] = c


def build_plan(shape: NetShape, fold_feature: bool = True) -> Plan:
    """Static execution plan of one network.

    ``fold_feature`` (view-direction nets only): ``feature_linear`` has no activation (run_nerf_helpers.py:126), so
    views(feature(h)) = relu([W_v1 W_f] h + W_vd dir + [W_v1 b_f + b_v]) with W_v1 = views weight[:, :W].  The chain
    then skips the 256x256 feature step in the forward pass, one of the two dgrad steps behind it and one 256x256
    wgrad item (with 4 + 4 slabs of stash per tile): M = W_v1 W_f and b' are rebuilt by ``dln_mlp_fold`` whenever
    the weights change, and the gradient of M is unfolded into dW_v1 = dM W_f^T + db' b_f^T, dW_f = W_v1^T dM, db_f = W_v1^T db'
    by ``dln_mlp_unfold_grads`` -- identical in exact arithmetic, one bf16 rounding fewer in practice."""
    shape.validate()
    D, W = shape.D, shape.W
    # netwidth < 256: every layer still runs as a 256- (views: 128-) column step of the kernels with the missing weight
    # rows / K columns packed as zeros; `n_valid32` tells the epilogues which columns are real.  The folded feature
    # layer's kernels are 256-wide, so narrow nets run the unfolded plan.
    fold_feature = fold_feature and W == 256
    NV, NVH = W // 32, W // 64            # valid output columns / 32 of a hidden (W) and of the views (W/2) layer

    def k_slabs(width, first=0):
        """(slab list, K=16-step counts) covering `width` input features held in slabs first, first+1, ..."""
        n = _ceil_div(width, 64)
        return [first + j for j in range(n)], [min(4, _ceil_div(width - 64 * j, 16)) for j in range(n)]

    pl = Plan(shape=shape)
    off = 0
    for name, shp in shape.param_shapes():
        pl.offsets[name] = off
        n = 1
        for d in shp:
            n *= d
        off += (n + 3) // 4 * 4      # keep every tensor 16-byte aligned inside the flat buffer
    pl.n_params = off
    pl.fold = bool(fold_feature and shape.use_viewdirs)
    pl.n_flat = off
    if pl.fold:
        pl.off_M, pl.off_bM = off, off + (W // 2) * W
        pl.n_flat = pl.off_bM + W // 2
    O = pl.offsets
    if shape.sem_K:
        K = shape.sem_K
        so = L.SemOffsets()
        so.w_f, so.b_f = O["feature_linear.weight"], O["feature_linear.bias"]
        so.w_s1, so.b_s1 = O["semantic_linear.0.weight"], O["semantic_linear.0.bias"]
        so.w_s2, so.b_s2 = O["semantic_linear.1.weight"], O["semantic_linear.1.bias"]
        so.A = pl.n_flat
        so.a = so.A + (W // 2) * W
        so.Sw = so.a + W // 2
        so.sc = so.Sw + K * W
        so.K = K
        pl.n_flat = so.sc + (K + 3) // 4 * 4
        pl.sem = so
    kc_pts = _ceil_div(shape.input_ch, 16)
    kc_dir = _ceil_div(shape.input_ch_views, 16)

    # ------------------------------------------------------------------ forward chain
    fwd = L.ChainProgram()
    fwd.backward, fwd.use_viewdirs, fwd.out_ch = 0, int(shape.use_viewdirs), shape.out_ch
    fwd.L_pts, fwd.L_dir = shape.L_pts, shape.L_dir
    blob = 0
    steps = 0

    def add_job(jobs, name, ld, row0, col0, n_valid, k_valid, transposed, n_rows, dst):
        src = O[name] if isinstance(name, str) else int(name)          # parameter name or raw float offset
        jobs.append(L.PackJob(src, ld, row0, col0, n_valid, k_valid, transposed, n_rows, dst))

    H_slot = lambda i: 2 + 4 * i          # forward stash slot of the output of pts layer i
    for i in range(D):
        st = fwd.steps[steps]
        wname = "pts_linears.%d.weight" % i
        ld = shape.fan_in(i)
        st.w_off, st.bias_off, st.n_out, st.n_valid32 = blob, O["pts_linears.%d.bias" % i], 256, NV
        hs, hc = k_slabs(W)
        hcols = [(64 * j, min(64, W - 64 * j)) for j in range(len(hs))]
        if i == 0:
            slabs, cnts, cols = [4], [kc_pts], [(0, shape.input_ch)]
        elif (i - 1) in shape.skips:
            slabs = [4] + hs
            cnts = [kc_pts] + hc
            cols = [(0, shape.input_ch)] + [(shape.input_ch + c, kv) for c, kv in hcols]
        else:
            slabs, cnts, cols = hs, hc, hcols
        _set_k(st, slabs, cnts)
        for (c0, kv) in cols:
            add_job(pl.fwd_jobs, wname, ld, 0, c0, W, kv, 0, 256, blob)
            blob += 256 * 128
        st.stash_slot, st.mask_slot = H_slot(i), i
        st.epi = L.EPI_RELU
        if i == D - 1:
            if shape.use_viewdirs:
                st.epi, st.n_heads = L.EPI_RELU_SIGMA, 1
                st.head_off, st.head_bias_off = O["alpha_linear.weight"], O["alpha_linear.bias"]
            else:
                st.epi, st.n_heads = L.EPI_RELU_OUT, shape.output_ch
                st.head_off, st.head_bias_off = O["output_linear.weight"], O["output_linear.bias"]
        steps += 1
    feat_slot = H_slot(D)
    pl.h_last_slot = H_slot(D - 1)
    hv_slot = feat_slot + (0 if pl.fold else 4)
    if shape.use_viewdirs:
        if not pl.fold:
            st = fwd.steps[steps]
            st.w_off, st.bias_off, st.n_out, st.epi, st.n_valid32 = blob, O["feature_linear.bias"], 256, L.EPI_LINEAR, NV
            hs, hc = k_slabs(W)
            _set_k(st, hs, hc)
            for j in hs:
                add_job(pl.fwd_jobs, "feature_linear.weight", W, 0, 64 * j, W, min(64, W - 64 * j), 0, 256, blob)
                blob += 256 * 128
            st.stash_slot, st.mask_slot = feat_slot, -1
            steps += 1
        st = fwd.steps[steps]
        st.w_off, st.n_out, st.epi, st.n_valid32 = blob, 128, L.EPI_RELU_RGB, NVH
        st.bias_off = pl.off_bM if pl.fold else O["views_linears.0.bias"]
        hs, hc = k_slabs(W)
        _set_k(st, hs + [4], hc + [kc_dir])     # slab 4 holds the encoded direction by now
        ldv = W + shape.input_ch_views
        for j in hs:
            if pl.fold:      # K slabs of M = W_v1 W_f act directly on the last hidden layer
                add_job(pl.fwd_jobs, pl.off_M, W, 0, 64 * j, 128, 64, 0, 128, blob)
            else:
                add_job(pl.fwd_jobs, "views_linears.0.weight", ldv, 0, 64 * j, W // 2, min(64, W - 64 * j), 0, 128, blob)
            blob += 128 * 128
        add_job(pl.fwd_jobs, "views_linears.0.weight", ldv, 0, W, W // 2, shape.input_ch_views, 0, 128, blob)
        blob += 128 * 128
        st.n_heads, st.head_off, st.head_bias_off = 3, O["rgb_linear.weight"], O["rgb_linear.bias"]
        st.stash_slot, st.mask_slot = hv_slot, D
        steps += 1
        pl.fwd_slots = hv_slot + 2
        pl.mask_slots = D + 1
    else:
        pl.fwd_slots = feat_slot
        pl.mask_slots = D
    # slab 4 carries the encoded position until the last pts layer that reads it, then the encoded direction
    fwd.reload_step = max([0] + [i for i in range(D) if (i - 1) in shape.skips]) if shape.use_viewdirs else -1
    fwd.n_steps, fwd.stash_slots, fwd.mask_slots = steps, pl.fwd_slots, pl.mask_slots
    pl.fwd, pl.fwd_blob_bytes = fwd, blob

    # ------------------------------------------------------------------ dgrad chain
    bwd = L.ChainProgram()
    bwd.backward, bwd.use_viewdirs, bwd.out_ch = 1, int(shape.use_viewdirs), shape.out_ch
    bwd.L_pts, bwd.L_dir = shape.L_pts, shape.L_dir
    blob = 0
    steps = 0
    bwd.pro_slot = 1
    bwd.reload_step = -1
    bwd.pro_valid = (W // 2 if shape.use_viewdirs else W) if W != 256 else 0
    if shape.use_viewdirs:
        bwd.pro_head_off, bwd.pro_mask_slot = O["rgb_linear.weight"], D
        dzv_slot, dzf_slot = 1, 3
        first_dz = 3 if pl.fold else 7
        dz_slot = lambda l: first_dz + 4 * (D - 1 - l)
        ldv = W + shape.input_ch_views
        if pl.fold:
            st = bwd.steps[steps]        # dH_{D-1} = dZ_v * M + d sigma * w_alpha ; mask
            st.w_off, st.n_out, st.epi = blob, 256, L.EPI_BWD_MASK_SIGMA
            _set_k(st, [0, 1], [4, 4])
            for j in range(2):
                add_job(pl.bwd_jobs, pl.off_M, W, 64 * j, 0, 256, 64, 1, 256, blob)
                blob += 256 * 128
            st.n_heads, st.head_off = 1, O["alpha_linear.weight"]
            st.stash_slot, st.mask_slot = dz_slot(D - 1), D - 1
            steps += 1
        else:
            st = bwd.steps[steps]            # d feature = dZ_v * W_views[:, :W]
            st.w_off, st.n_out, st.epi, st.n_valid32 = blob, 256, L.EPI_BWD_COPY, NV
            vs, vc = k_slabs(W // 2)
            _set_k(st, vs, vc)
            for j in vs:
                add_job(pl.bwd_jobs, "views_linears.0.weight", ldv, 64 * j, 0, W, min(64, W // 2 - 64 * j), 1, 256, blob)
                blob += 256 * 128
            st.stash_slot, st.mask_slot = dzf_slot, -1
            steps += 1
            st = bwd.steps[steps]            # dH_{D-1} = d feature * W_feature + d sigma * w_alpha ; mask
            st.w_off, st.n_out, st.epi, st.n_valid32 = blob, 256, L.EPI_BWD_MASK_SIGMA, NV
            hs, hc = k_slabs(W)
            _set_k(st, hs, hc)
            for j in hs:
                add_job(pl.bwd_jobs, "feature_linear.weight", W, 64 * j, 0, W, min(64, W - 64 * j), 1, 256, blob)
                blob += 256 * 128
            st.n_heads, st.head_off = 1, O["alpha_linear.weight"]
            st.stash_slot, st.mask_slot = dz_slot(D - 1), D - 1
            steps += 1
    else:
        bwd.pro_head_off, bwd.pro_mask_slot = O["output_linear.weight"], D - 1
        dz_slot = lambda l: 1 + 4 * (D - 1 - l)
    for l in range(D - 1, 0, -1):        # dH_{l-1} = dZ_l * W_l[:, h-part] ; mask_{l-1}
        st = bwd.steps[steps]
        st.w_off, st.n_out, st.epi, st.n_valid32 = blob, 256, L.EPI_BWD_MASK, NV
        hs, hc = k_slabs(W)
        _set_k(st, hs, hc)
        c0 = shape.input_ch if (l - 1) in shape.skips else 0
        for j in hs:
            add_job(pl.bwd_jobs, "pts_linears.%d.weight" % l, shape.fan_in(l), 64 * j, c0, W, min(64, W - 64 * j), 1, 256, blob)
            blob += 256 * 128
        st.stash_slot, st.mask_slot = dz_slot(l - 1), l - 1
        steps += 1
    if steps == 0:
        raise NotImplementedError("netdepth=1 without view directions has no hidden dgrad step")
    pl.bwd_slots = dz_slot(0) + 4
    bwd.n_steps, bwd.stash_slots, bwd.mask_slots = steps, pl.bwd_slots, pl.mask_slots
    pl.bwd, pl.bwd_blob_bytes = bwd, blob

    # ------------------------------------------------------------------ wgrad items
    def item(a_slot, a_n, b_slot, b_n, wname, ld, col_off, n_cols, row_off, n_rows, bname=None, db_col=0, db_n=0):
        it = L.WgradItem()
        it.a_bwd_stash, it.a_slot, it.a_nslab = 1, a_slot, a_n
        it.b_from_bwd, it.b_slot, it.b_nslab = 0, b_slot, b_n
        it.dw_off, it.ld, it.col_off, it.n_cols = (O[wname] if isinstance(wname, str) else int(wname)), ld, col_off, n_cols
        it.row_off, it.n_rows = row_off, n_rows
        it.db_off = -1 if bname is None else (O[bname] if isinstance(bname, str) else int(bname))
        it.db_col_off, it.db_n = db_col, db_n
        pl.wgrad.append(it)

    # (the slab counts stay 4 / 2: the padded slabs of a narrow net hold zeros; rows / columns are limited to W)
    for l in range(D):
        wn, bn, ld = "pts_linears.%d.weight" % l, "pts_linears.%d.bias" % l, shape.fan_in(l)
        if l == 0:
            item(dz_slot(0), 4, 0, 1, wn, ld, 0, shape.input_ch, 0, W, bn, 0, W)
        elif (l - 1) in shape.skips:
            item(dz_slot(l), 4, 0, 1, wn, ld, 0, shape.input_ch, 0, W, bn, 0, W)
            item(dz_slot(l), 4, H_slot(l - 1), 4, wn, ld, shape.input_ch, W, 0, W)
        else:
            item(dz_slot(l), 4, H_slot(l - 1), 4, wn, ld, 0, W, 0, W, bn, 0, W)
    if shape.use_viewdirs:
        if not pl.fold:
            item(dzf_slot, 4, H_slot(D - 1), 4, "feature_linear.weight", W, 0, W, 0, W, "feature_linear.bias", 0, W)
        item(0, 1, H_slot(D - 1), 4, "alpha_linear.weight", W, 0, W, 3, 1, "alpha_linear.bias", 3, 1)
        ldv = W + shape.input_ch_views
        if pl.fold:      # dM = dZ_v^T H_{D-1} and db' = colsum(dZ_v) into the scratch tail of the gradient buffer
            item(dzv_slot, 2, H_slot(D - 1), 4, pl.off_M, W, 0, 256, 0, 128, pl.off_bM, 0, 128)
        else:
            item(dzv_slot, 2, feat_slot, 4, "views_linears.0.weight", ldv, 0, W, 0, W // 2, "views_linears.0.bias", 0, W // 2)
        item(dzv_slot, 2, 1, 1, "views_linears.0.weight", ldv, W, shape.input_ch_views, 0, W // 2)
        item(0, 1, hv_slot, 2, "rgb_linear.weight", W // 2, 0, W // 2, 0, 3, "rgb_linear.bias", 0, 3)
    else:
        item(0, 1, H_slot(D - 1), 4, "output_linear.weight", W, 0, W, 0, shape.output_ch,
             "output_linear.bias", 0, shape.output_ch)
    return pl
