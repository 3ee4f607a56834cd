"""Pure-numpy interpreter of the MLP execution plan (depth-lidar-nerf_b200/plan.py): executes the chain
programs, pack jobs and wgrad items with the semantics the CUDA kernels implement (csrc/mlp_kernels.cu),
in float64 and without the bf16 rounding or the smem swizzle.  Test infrastructure: lets the CPU suite
prove that the *plan* (offsets, transposes, slab/slot maps, K-slab lists, heads) reproduces the
reference MLP and its gradients before any GPU time is spent."""
from __future__ import annotations

import numpy as np

import dlnerf_b200 as dn

L = dn._lib


def _stage_matrix(flat, job):
    """Logical [n_rows, 64] weight stage a PackJob describes."""
    M = np.zeros((job.n_rows, 64))
    W = flat[job.src_off:]
    for n in range(min(job.n_rows, job.n_valid)):
        for k in range(job.k_valid):
            if job.transposed:
                M[n, k] = W[(job.row0 + k) * job.ld + job.col0 + n]
            else:
                M[n, k] = W[(job.row0 + n) * job.ld + job.col0 + k]
    return M


def _stage_matrix_fast(flat, job):
    M = np.zeros((job.n_rows, 64))
    nv, kv = min(job.n_rows, job.n_valid), job.k_valid
    if job.transposed:
        rows = job.row0 + np.arange(kv)
        cols = job.col0 + np.arange(nv)
        M[:nv, :kv] = flat[job.src_off + rows[None, :] * job.ld + cols[:, None]]
    else:
        rows = job.row0 + np.arange(nv)
        cols = job.col0 + np.arange(kv)
        M[:nv, :kv] = flat[job.src_off + rows[:, None] * job.ld + cols[None, :]]
    return M


def _stages_by_offset(flat, jobs):
    return {j.dst_off: _stage_matrix_fast(flat, j) for j in jobs}


def extend_flat(plan, flat):
    """[parameters | M = W_v1 W_f | b' = W_v1 b_f + b_v]: what dln_mlp_fold appends behind the parameters."""
    out = np.zeros(plan.n_flat)
    out[:plan.n_params] = flat[:plan.n_params]
    if plan.fold:
        O, W = plan.offsets, plan.shape.W
        ldv = W + plan.shape.input_ch_views
        Wv1 = flat[O["views_linears.0.weight"]: O["views_linears.0.weight"] + (W // 2) * ldv].reshape(W // 2, ldv)[:, :W]
        Wf = flat[O["feature_linear.weight"]: O["feature_linear.weight"] + W * W].reshape(W, W)
        bf = flat[O["feature_linear.bias"]: O["feature_linear.bias"] + W]
        bv = flat[O["views_linears.0.bias"]: O["views_linears.0.bias"] + W // 2]
        out[plan.off_M: plan.off_M + (W // 2) * W] = (Wv1 @ Wf).reshape(-1)
        out[plan.off_bM: plan.off_bM + W // 2] = Wv1 @ bf + bv
    return out


def unfold_grads(plan, flat, g):
    """dln_mlp_unfold_grads: scratch (dM, db') -> dW_v1 += dM W_f^T + db' b_f^T (feature = W_f h + b_f),
    dW_f += W_v1^T dM, db_f += W_v1^T db', db_v += db'."""
    if not plan.fold:
        return g
    O, W = plan.offsets, plan.shape.W
    ldv = W + plan.shape.input_ch_views
    Wv = flat[O["views_linears.0.weight"]: O["views_linears.0.weight"] + (W // 2) * ldv].reshape(W // 2, ldv)
    Wf = flat[O["feature_linear.weight"]: O["feature_linear.weight"] + W * W].reshape(W, W)
    dM = g[plan.off_M: plan.off_M + (W // 2) * W].reshape(W // 2, W)
    db = g[plan.off_bM: plan.off_bM + W // 2]
    gWv = g[O["views_linears.0.weight"]: O["views_linears.0.weight"] + (W // 2) * ldv].reshape(W // 2, ldv)
    bf = flat[O["feature_linear.bias"]: O["feature_linear.bias"] + W]
    gWv[:, :W] += dM @ Wf.T + np.outer(db, bf)
    g[O["feature_linear.weight"]: O["feature_linear.weight"] + W * W] += (Wv[:, :W].T @ dM).reshape(-1)
    g[O["feature_linear.bias"]: O["feature_linear.bias"] + W] += Wv[:, :W].T @ db
    g[O["views_linears.0.bias"]: O["views_linears.0.bias"] + W // 2] += db
    return g


def run_forward(plan, flat, enc_pts, enc_dir):
    """enc_pts [P, <=64], enc_dir [P, <=64] (already encoded rows).  Returns (out [P,out_ch], stash dict
    slot -> [P,64], masks dict slot -> bool [P, n_out])."""
    prog = plan.fwd
    P = enc_pts.shape[0]
    slabs = [np.zeros((P, 64)) for _ in range(5)]
    slabs[4][:, :enc_pts.shape[1]] = enc_pts
    stash = {0: slabs[4].copy()}
    masks = {}
    stages = _stages_by_offset(flat, plan.fwd_jobs)
    sigma = None
    out = None
    for s in range(prog.n_steps):
        st = prog.steps[s]
        acc = np.zeros((P, st.n_out))
        for j in range(st.nk):
            Wst = stages[st.w_off + j * st.n_out * 128]
            k = 16 * st.kcnt[j]
            acc += slabs[st.kslab[j]][:, :k] @ Wst[:, :k].T
        nv = 32 * st.n_valid32 if st.n_valid32 else st.n_out      # columns beyond are padding: zeros, no bias, no head
        x = np.zeros((P, st.n_out))
        x[:, :nv] = acc[:, :nv] + flat[st.bias_off: st.bias_off + nv]
        if st.epi in (L.EPI_RELU, L.EPI_RELU_SIGMA, L.EPI_RELU_RGB, L.EPI_RELU_OUT):
            if st.mask_slot >= 0:
                masks[st.mask_slot] = x > 0
            x = np.maximum(x, 0)
        heads = None
        if st.n_heads:
            Hw = flat[st.head_off: st.head_off + st.n_heads * nv].reshape(st.n_heads, nv)
            heads = x[:, :nv] @ Hw.T + flat[st.head_bias_off: st.head_bias_off + st.n_heads]
        for i in range(st.n_out // 64):
            slabs[i] = x[:, 64 * i: 64 * i + 64].copy()
            if st.stash_slot >= 0:
                stash[st.stash_slot + i] = slabs[i].copy()
        if s == prog.reload_step:          # slab 4: encoded position -> encoded direction
            slabs[4] = np.zeros((P, 64))
            slabs[4][:, :enc_dir.shape[1]] = enc_dir
            stash[1] = slabs[4].copy()
        if st.epi == L.EPI_RELU_SIGMA:
            sigma = heads[:, 0]
        elif st.epi == L.EPI_RELU_RGB:
            out = np.concatenate([heads[:, :3], sigma[:, None]], 1)
        elif st.epi == L.EPI_RELU_OUT:
            out = heads[:, :prog.out_ch]
    return out, stash, masks


def run_backward(plan, flat, d_out, masks):
    """Returns the backward stash dict slot -> [P,64]."""
    prog = plan.bwd
    P = d_out.shape[0]
    slabs = [np.zeros((P, 64)) for _ in range(5)]
    stash = {}
    nh = 3 if prog.use_viewdirs else prog.out_ch
    width = 128 if prog.use_viewdirs else 256
    pv = prog.pro_valid if prog.pro_valid > 0 else width
    Hw = flat[prog.pro_head_off: prog.pro_head_off + nh * pv].reshape(nh, pv)
    dz = np.zeros((P, width))
    dz[:, :pv] = (d_out[:, :nh] @ Hw) * masks[prog.pro_mask_slot][:, :pv]
    for i in range(width // 64):
        slabs[i] = dz[:, 64 * i: 64 * i + 64].copy()
        stash[prog.pro_slot + i] = slabs[i].copy()
    slabs[4][:, :prog.out_ch] = d_out[:, :prog.out_ch]
    stash[0] = slabs[4].copy()
    dsig = d_out[:, 3] if prog.out_ch > 3 else np.zeros(P)
    stages = _stages_by_offset(flat, plan.bwd_jobs)
    for s in range(prog.n_steps):
        st = prog.steps[s]
        acc = np.zeros((P, st.n_out))
        for j in range(st.nk):
            Wst = stages[st.w_off + j * st.n_out * 128]
            k = 16 * st.kcnt[j]
            acc += slabs[st.kslab[j]][:, :k] @ Wst[:, :k].T
        nv = 32 * st.n_valid32 if st.n_valid32 else st.n_out
        x = np.zeros((P, st.n_out))
        x[:, :nv] = acc[:, :nv]
        if st.epi == L.EPI_BWD_MASK_SIGMA:
            x[:, :nv] += dsig[:, None] * flat[st.head_off: st.head_off + nv][None, :]
        if st.epi in (L.EPI_BWD_MASK, L.EPI_BWD_MASK_SIGMA):
            x = x * masks[st.mask_slot]
        for i in range(st.n_out // 64):
            slabs[i] = x[:, 64 * i: 64 * i + 64].copy()
            if st.stash_slot >= 0:
                stash[st.stash_slot + i] = slabs[i].copy()
    return stash


def run_wgrad(plan, stash_f, stash_b):
    g = np.zeros(plan.n_flat)
    for it in plan.wgrad:
        A = np.concatenate([stash_b[it.a_slot + i] for i in range(it.a_nslab)], 1)        # [P, 64*a_nslab]
        src = stash_b if it.b_from_bwd else stash_f
        B = np.concatenate([src[it.b_slot + i] for i in range(it.b_nslab)], 1)            # [P, 64*b_nslab]
        acc = A.T @ B
        for i in range(it.n_rows):
            row = it.row_off + i
            g[it.dw_off + i * it.ld + it.col_off: it.dw_off + i * it.ld + it.col_off + it.n_cols] += acc[row, :it.n_cols]
        if it.db_off >= 0:
            colsum = A.sum(0)
            g[it.db_off: it.db_off + it.db_n] += colsum[it.db_col_off: it.db_col_off + it.db_n]
    return g
