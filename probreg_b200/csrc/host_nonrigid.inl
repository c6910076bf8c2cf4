// host_nonrigid.inl: the non-rigid entry points of include/cpd_b200.h (dense G, low-rank factors, correspondence priors) -- part of the single translation unit cpd_b200.cu (included at its end; uses its handle type, error
// macros and helpers).  Split out for readability only.
// ---------------------------------------------------------------------------------------------
// non-rigid CPD with a dense G, resident on the device
// ---------------------------------------------------------------------------------------------
namespace {
// state shared by the dense and the low-rank non-rigid loops: W = 0 (cpd.py:281), T = Y, sigma2, w
int nonrigid_common_begin(cpd_ctx* h, double lmd, double sigma2, double w) {
    const long long m = h->m;
    if (h->nr_m != m) {
        TRY(dev_alloc(&h->d_W, (size_t)m * 3));
        TRY(dev_alloc(&h->d_B, (size_t)m * 3));
        TRY(dev_alloc(&h->d_ts2, (size_t)m * 3));
        TRY(dev_alloc(&h->d_wgt, (size_t)m));
        TRY(dev_alloc(&h->d_nrpart, (size_t)blocks_for(m) * 2));
        TRY(dev_alloc(&h->d_ipiv, (size_t)m));
        TRY(dev_alloc(&h->d_info, 2));          // [0] cuSOLVER's info, [1] the rank lr_pchol_kernel found
        if (h->d_G) { cudaFree(h->d_G); h->d_G = nullptr; }
        if (h->d_A) { cudaFree(h->d_A); h->d_A = nullptr; }
        h->nr_m = m;
        h->work_dev = 0;
    }
    h->prior_on = false;                         // priors are set after begin (cpd_nonrigid_set_prior)
    if (!h->sol) {
        SOLV(g_sol.Create(&h->sol));
        SOLV(g_sol.SetStream(h->sol, h->stream));
        SOLV(g_sol.CreateParams(&h->sol_params));
    }
    TRY(ensure_stats(h));
    const DevState& hs = h->h_state;
    CU(cudaMemsetAsync(h->d_W, 0, (size_t)m * 3 * sizeof(double), h->stream));                       // cpd.py:281
    h->lr_w_stale = false;
    CU(cudaMemsetAsync(h->d_info, 0, sizeof(int), h->stream));
    nr_identity_kernel<<<blocks_for(m), THREADS, 0, h->stream>>>(h->d_yc, hs.cy[0], hs.cy[1], hs.cy[2], m, h->d_ts);   // T = Y + G 0
    KCHECK();
    h->launches += 1;
    h->nr_lmd = lmd;
    h->h_state.sigma2 = sigma2;
    h->h_state.q = 0.0;
    h->h_state.w = w;
    h->h_state.tf_kind = CPD_TF_NONRIGID;
    h->h_state.err = 0;
    return upload_state(h);
}
// workspace of the LU of an n x n system stored at `a`
int solver_workspace(cpd_ctx* h, long long n, double* a) {
    size_t wd = 0, wh = 0;
    SOLV(g_sol.XgetrfBuf(h->sol, h->sol_params, n, n, CUDA_R_64F_, a, n, CUDA_R_64F_, &wd, &wh));
    if (wd > h->work_dev || !h->d_work) {
        if (h->d_work) cudaFree(h->d_work);
        h->d_work = nullptr;
        CU(cudaMalloc(&h->d_work, std::max<size_t>(wd, 16)));
        h->work_dev = std::max<size_t>(wd, 16);
    }
    if (wh > h->work_host || !h->h_work) {
        free(h->h_work);
        h->h_work = malloc(std::max<size_t>(wh, 16));
        h->work_host = std::max<size_t>(wh, 16);
    }
    return CPD_OK;
}
// Orthonormalise the `rank` columns of X ([rank][ld]) in place, column by column: classical Gram-Schmidt, each column projected
// twice (6 launches per column; kept as the cross-check of the blocked version below: CPD_B200_LR_ORTH=columnwise)
int lr_orthonormalise_columnwise(cpd_ctx* h, double* X, int rank) {
    const long long m = h->m, ld = h->mpad;
    const unsigned nb = blocks_for(m);
    const int stride = rank + 1;
    long long per_slice = 8192;                              // points per slice of a dot product (tests lower it to reach nsl > 1)
    if (const char* e = getenv("CPD_B200_LR_SLICE_POINTS")) per_slice = std::max<long long>(32, atoll(e));
    const unsigned nsl = (unsigned)std::min<long long>(LR_SLICES, std::max<long long>(1, (m + per_slice - 1) / per_slice));
    const size_t round_sz = (size_t)LR_SLICES * stride;     // d_lr_coef: [3 rounds][LR_SLICES][rank + 1] slice partials
    for (int j = 0; j < rank; ++j) {
        for (int round = 0; round < 2; ++round) {
            double* part = h->d_lr_coef + round * round_sz;
            lr_dots_kernel<<<dim3((unsigned)(j + 1), nsl), THREADS, 0, h->stream>>>(X, m, ld, j, 0, stride, part);
            if (j > 0) lr_project_kernel<<<nb, THREADS, 0, h->stream>>>(X, m, ld, j, stride, (int)nsl, part);
            h->launches += j > 0 ? 2 : 1;
        }
        double* part2 = h->d_lr_coef + 2 * round_sz;
        lr_dots_kernel<<<dim3(1, nsl), THREADS, 0, h->stream>>>(X, m, ld, j, j, stride, part2);
        lr_scale_kernel<<<nb, THREADS, 0, h->stream>>>(X, m, ld, j, stride, (int)nsl, h->d_lr_coef, part2);
        h->launches += 2;
    }
    KCHECK();
    return CPD_OK;
}
// dst[c][i] = sum_j G_ij src[c][j] for all `rank` columns.  The product shards over rows with one exchange: in a multi-rank
// handle (sources replicated, identically ordered on every rank) each rank forms its contiguous share of the rows into a zeroed
// buffer and one all-reduce -- a sum of one value and zeros, hence exact and identical everywhere -- gathers them.
// The rows are formed on the tensor cores (gram_umma.cuh: tcgen05, TF32 x 3); CPD_B200_LR_GRAM=simt selects the CUDA-core
// kernel the tensor-core path is checked against on its first use in a process (and which the CPU emulation build uses).
#ifndef CPD_HOST_EMU
int lr_gram_rows_umma(cpd_ctx* h, const double* src, double* dst, int rank, long long i_lo, long long i_hi) {
    const long long ld = h->mpad, rows = i_hi - i_lo;
    long long chunk = 1024;                                   // points per FP32 TMEM accumulation (error ~7e-9 per point, see DESIGN)
    if (const char* e = getenv("CPD_B200_LR_CHUNK")) chunk = std::max<long long>(GU_KS, atoll(e) / GU_KS * GU_KS);
    chunk = std::max(chunk, (ld / 128 + GU_KS - 1) / GU_KS * GU_KS);                // at most 128 chunk partials
    const int nq = (int)((ld + chunk - 1) / chunk);
    const int ntiles = (int)((rows + GU_ROWS - 1) / GU_ROWS);
    const long long ldp = (long long)ntiles * GU_ROWS;
    static bool attr_set = false;
    if (!attr_set) {
        CU(cudaFuncSetAttribute(gu_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GU_SMEM));
        attr_set = true;
    }
    for (int c0 = 0; c0 < rank; c0 += GU_NMAX) {
        const int nc = std::min(GU_NMAX, rank - c0), n16 = (nc + 15) / 16 * 16;
        const size_t need_planes = (size_t)2 * n16 * ld, need_part = (size_t)nq * n16 * ldp;
        if (need_planes > h->gu_planes_cap) { TRY(dev_alloc(&h->d_gu_planes, need_planes)); h->gu_planes_cap = need_planes; }
        if (need_part > h->gu_part_cap) { TRY(dev_alloc(&h->d_gu_part, need_part)); h->gu_part_cap = need_part; }
        gu_split_kernel<<<dim3(blocks_for(ld), (unsigned)n16), THREADS, 0, h->stream>>>(src + (size_t)c0 * ld, h->m, ld, nc, n16, ld, h->d_gu_planes);
        CUtensorMap map;
        if (gu_make_map(&map, h->d_gu_planes, ld, 2 * n16, n16) != 0) return fail(CPD_ERR_CUDA, "cuTensorMapEncodeTiled failed");
        gu_gram_kernel<<<h->sm_count, GU_THREADS, GU_SMEM, h->stream>>>(map, h->d_lr_pts, ld, (int)chunk, i_lo, i_hi, n16, h->d_gu_part, ldp);
        gu_reduce_kernel<<<dim3((unsigned)((rows + 4 * THREADS - 1) / (4 * THREADS)), (unsigned)nc), THREADS, 0, h->stream>>>(
            h->d_gu_part, nq, n16, ldp, nc, rows, i_lo, ld, dst + (size_t)c0 * ld);
        KCHECK();
        h->launches += 3;
    }
    return CPD_OK;
}
// exact integer-digit product (gram_i8.cuh): the default
int lr_gram_rows_i8(cpd_ctx* h, const double* src, double* dst, int rank, long long i_lo, long long i_hi, bool a_in_tmem) {
    const long long ld = h->mpad, rows = i_hi - i_lo;
    const long long chunk = std::min<long long>(GI_MAX_CHUNK, ld);          // ld is a multiple of 512, hence of GI_KS
    const int nq = (int)((ld + chunk - 1) / chunk);
    const int ntiles = (int)((rows + GI_ROWS - 1) / GI_ROWS);
    const long long ldp = (long long)ntiles * GI_ROWS;
    static bool attr_set = false;
    if (!attr_set) {
        CU(cudaFuncSetAttribute(gi_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GI_SMEM));
        CU(cudaFuncSetAttribute(gi_gram_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GI_TS_SMEM));
        attr_set = true;
    }
    const int passes = (rank + GI_NMAX - 1) / GI_NMAX, per = (rank + passes - 1) / passes;
    const size_t ncm = (size_t)rank + GI_NMAX + 16;
    if (ncm > h->gi_colmax_cap) { TRY(dev_alloc(&h->d_gi_colmax, ncm)); h->gi_colmax_cap = ncm; }
    CU(cudaMemsetAsync(h->d_gi_colmax, 0, ncm * sizeof(double), h->stream));
    gi_colmax_kernel<<<(unsigned)rank, THREADS, 0, h->stream>>>(src, h->m, ld, h->d_gi_colmax);
    if ((size_t)ld > h->gi_pairs_cap) { TRY(dev_alloc(&h->d_gi_pairs, (size_t)ld)); h->gi_pairs_cap = (size_t)ld; }
    gi_pairs_kernel<<<blocks_for(ld / 2), THREADS, 0, h->stream>>>(h->d_lr_pts, ld / 2, h->d_gi_pairs);
    h->launches += 2;
    for (int c0 = 0; c0 < rank; c0 += per) {
        const int nc = std::min(per, rank - c0), n16 = (nc + 15) / 16 * 16;
        const size_t need_planes = (size_t)(ld / GI_KS) * 3 * GI_PLANE, need_part = (size_t)nq * n16 * ldp;
        if (need_planes > h->gi_planes_cap) { TRY(dev_alloc(&h->d_gi_planes, need_planes)); h->gi_planes_cap = need_planes; }
        if (need_part > h->gi_part_cap) { TRY(dev_alloc(&h->d_gi_part, need_part)); h->gi_part_cap = need_part; }
        gi_split_kernel<<<dim3(blocks_for(ld / 16), (unsigned)n16), THREADS, 0, h->stream>>>(src + (size_t)c0 * ld, h->m, ld, nc, n16, ld,
                                                                                          h->d_gi_colmax + c0, h->d_gi_planes);
        if (a_in_tmem)
            gi_gram_ts_kernel<<<h->sm_count, GI_TS_THREADS, GI_TS_SMEM, h->stream>>>(h->d_gi_planes, h->d_lr_pts, h->d_gi_pairs, ld, (int)chunk, i_lo, i_hi,
                                                                                 n16, h->d_gi_colmax + c0, h->d_gi_part, ldp);
        else
            gi_gram_kernel<<<h->sm_count, GI_THREADS, GI_SMEM, h->stream>>>(h->d_gi_planes, h->d_lr_pts, h->d_gi_pairs, ld, (int)chunk, i_lo, i_hi, n16,
                                                                           h->d_gi_colmax + c0, h->d_gi_part, ldp);
        gi_reduce_kernel<<<dim3(blocks_for(rows), (unsigned)nc), THREADS, 0, h->stream>>>(h->d_gi_part, nq, n16, ldp, nc, rows, i_lo, ld,
                                                                                           dst + (size_t)c0 * ld);
        KCHECK();
        h->launches += 3;
    }
    return CPD_OK;
}
#endif
int lr_gram_rows_simt(cpd_ctx* h, const double* src, double* dst, int rank, long long i_lo, long long i_hi) {
    dim3 grid(blocks_for(i_hi - i_lo), (unsigned)((rank + LR_COLS - 1) / LR_COLS));
    lr_gram_apply_kernel<<<grid, THREADS, 0, h->stream>>>(h->d_lr_pts, h->m, h->mpad, src, h->mpad, rank, dst, i_lo, i_hi);
    KCHECK();
    h->launches += 1;
    return CPD_OK;
}
int lr_gram_apply(cpd_ctx* h, const double* src, double* dst, int rank) {
    long long i_lo = 0, i_hi = h->m;
    const bool shard = h->comm != nullptr && h->world > 1;
    if (shard) {
        i_lo = h->m * h->rank / h->world;
        i_hi = h->m * (h->rank + 1) / h->world;
        CU(cudaMemsetAsync(dst, 0, (size_t)rank * h->mpad * sizeof(double), h->stream));
    }
    if (i_hi > i_lo) {
#ifdef CPD_HOST_EMU
        TRY(lr_gram_rows_simt(h, src, dst, rank, i_lo, i_hi));
#else
        // 0: tensor cores, exact integer digits, A operand in tensor memory (default); 3: the same with the A operand in shared memory;
        // 1: CUDA cores (FP32); 2: tensor cores, TF32 x 3 (FP32 TMEM accumulation: ~1e-5 relative, measured -- kept for comparison)
        static int mode = -1;
        static bool checked = false;
        if (mode < 0) {
            const char* e = getenv("CPD_B200_LR_GRAM");
            mode = (e && !strcmp(e, "simt")) ? 1 : ((e && !strcmp(e, "tf32")) ? 2 : ((e && !strcmp(e, "i8ss")) ? 3 : 0));   // 3: i8, A operand in smem
        }
        if (mode == 1) {
            TRY(lr_gram_rows_simt(h, src, dst, rank, i_lo, i_hi));
        } else {
            if (mode == 2) TRY(lr_gram_rows_umma(h, src, dst, rank, i_lo, i_hi));
            else TRY(lr_gram_rows_i8(h, src, dst, rank, i_lo, i_hi, mode != 3));
            if (!checked) {
                // first use in this process: the first rows of the first <= 16 columns once more on the CUDA cores.  A mismatch
                // is an error (a wrong descriptor or swizzle shows as O(1) differences), never a silent change of path.
                const long long rows = std::min<long long>(i_hi - i_lo, 256);
                const int cols = std::min(rank, LR_COLS);
                DevBuf<double> ref, res;
                TRY(ref.alloc((size_t)cols * h->mpad));
                TRY(res.alloc(2));
                TRY(lr_gram_rows_simt(h, src, ref.p, cols, i_lo, i_lo + rows));
                gu_compare_kernel<<<1, THREADS, 0, h->stream>>>(dst, ref.p, h->mpad, i_lo, rows, cols, res.p);
                KCHECK();
                double r[2] = {0.0, 0.0};
                CU(cudaMemcpyAsync(r, res.p, sizeof(r), cudaMemcpyDeviceToHost, h->stream));
                CU(cudaStreamSynchronize(h->stream));
                h->launches += 1;
                if (!(r[0] <= (mode == 2 ? 1e-4 : 5e-6) * r[1] + 1e-300))
                    return fail(CPD_ERR_CUDA, "tensor-core G X product disagrees with the CUDA-core kernel: max |diff| %.3e, max |value| %.3e", r[0], r[1]);
                checked = true;
            }
        }
#endif
    }
    if (shard) TRY(allreduce(h, dst, (size_t)rank * h->mpad));
    return CPD_OK;
}
// out[na][nb] = A diag(wt) Bm^T over the points.  The point range is cut into as many slices as the partial buffer holds (at most
// 256, at least 256 points each): many short CTAs instead of 8 long ones per tile; lr_merge_kernel adds the slices in a fixed order.
int lr_inner(cpd_ctx* h, const double* A, int na, long long lda, const double* Bm, int nb, long long ldb, const double* wt, double* out) {
    const int tiles = ((na + LR_TILE - 1) / LR_TILE) * ((nb + LR_TILE - 1) / LR_TILE);
    const long long by_cap = (long long)(h->lr_part_cap / ((size_t)na * nb));
    const int nsl = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(256, by_cap), h->m / 256));
    dim3 grid((unsigned)tiles, (unsigned)nsl);
    lr_inner_kernel<<<grid, THREADS, 0, h->stream>>>(A, na, lda, Bm, nb, ldb, wt, h->m, h->d_lr_part);
    lr_merge_kernel<<<blocks_for((long long)na * nb * 8), THREADS, 0, h->stream>>>(h->d_lr_part, nsl, na, nb, out);
    KCHECK();
    h->launches += 2;
    return CPD_OK;
}
// Symmetric form of the M-step's K x K system, once per set-up:  Bc ~= L L^T (pivoted Cholesky, lr_pchol_kernel),  Qt = Q L  (into
// d_lr_X, which the set-up no longer needs), so that G ~= Qt Qt^T and every M-step solves the symmetric positive definite
// (c I + Qt^T diag(p1) Qt) Z = Qt^T F in one CTA (lr_spd_solve_kernel) instead of an LU of (c I + Bc S) through cuSOLVER.
// cpd_nonrigid_lowrank_get hands out Q and Bc = L L^T (lr_llt_kernel): exactly the G of the iteration.
int lr_spd_form(cpd_ctx* h, int rank) {
    const long long m = h->m, ld = h->mpad;
    lr_pchol_kernel<<<1, LR_PCHOL_THREADS, 0, h->stream>>>(h->d_lr_Bc, rank, h->d_lr_Lt, h->d_info + 1);
    const unsigned nb = blocks_for(m);
    for (int j0 = 0; j0 < rank; j0 += LR_PANEL)
        lr_rotate_kernel<<<nb, THREADS, 0, h->stream>>>(h->d_lr_Q, m, ld, rank, h->d_lr_Lt, j0, std::min(LR_PANEL, rank - j0), h->d_lr_X);
    lr_llt_kernel<<<blocks_for((long long)rank * rank), THREADS, 0, h->stream>>>(h->d_lr_Lt, rank, h->d_lr_Bc);
    KCHECK();
    h->launches += 2 + (rank + LR_PANEL - 1) / LR_PANEL;
    return CPD_OK;
}
// out[n][n] = A diag(wt) Bm^T where the result is symmetric (S = Q^T diag(w) Q, Bc = Q^T (G Q)): 64 x 64 tiles of the lower triangle,
// slices of >= 256 points, as many as the partial buffer holds (at most 128).
int lr_inner_sym(cpd_ctx* h, const double* A, int n, long long lda, const double* Bm, long long ldb, const double* wt, double* out) {
    const int nt = (n + LRS_TILE - 1) / LRS_TILE, tiles = nt * (nt + 1) / 2;
    const long long by_cap = (long long)(h->lr_part_cap / ((size_t)n * n));
    const int nsl = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(128, by_cap), h->m / 256));
    dim3 grid((unsigned)tiles, (unsigned)nsl);
    lr_inner_sym_kernel<<<grid, THREADS, 0, h->stream>>>(A, n, lda, Bm, ldb, wt, h->m, h->d_lr_part);
    lr_merge_sym_kernel<<<blocks_for((long long)n * n * 8), THREADS, 0, h->stream>>>(h->d_lr_part, nsl, n, out);
    KCHECK();
    h->launches += 2;
    return CPD_OK;
}
// Orthonormalise the `rank` columns of X ([rank][ld]) in place: block Gram-Schmidt with re-orthogonalisation over panels of
// LR_PANEL columns (lowrank.cuh); numerically dependent columns become exactly zero.
int lr_orthonormalise(cpd_ctx* h, double* X, int rank) {
    if (const char* e = getenv("CPD_B200_LR_ORTH")) if (!strcmp(e, "columnwise")) return lr_orthonormalise_columnwise(h, X, rank);
    const long long m = h->m, ld = h->mpad;
    const unsigned nb = blocks_for(m);
    const int nblk = (int)std::min<long long>(LR_GRAM_BLOCKS, (m + LR_GRAM_PTS - 1) / LR_GRAM_PTS);
    double* C = h->d_lr_panel;                              // [rank][np] projection coefficients
    double* n0 = C + (size_t)LR_MAX_RANK * LR_PANEL;        // [LR_PANEL]: the norm^2 each column of the panel arrived with
    double* scale2 = n0 + LR_PANEL;                         // [LR_PANEL]: see lr_panel_chol_kernel
    double* T = scale2 + LR_PANEL;                          // [LR_PANEL][LR_PANEL]
    double* gpart = T + LR_PANEL * LR_PANEL;                // [nblk][LR_PANEL][LR_PANEL] block partials of the panel's Gram matrix
    for (int j0 = 0; j0 < rank; j0 += LR_PANEL) {
        const int np = std::min(LR_PANEL, rank - j0);
        double* P = X + (size_t)j0 * ld;
        lr_panel_gram_kernel<<<nblk, THREADS, 0, h->stream>>>(X, m, ld, j0, np, gpart);
        lr_panel_chol_kernel<<<1, LR_CHOL_THREADS, 0, h->stream>>>(gpart, nblk, np, n0, 2, scale2, T);
        h->launches += 2;
        for (int pass = 0; pass < 2; ++pass) {
            if (j0 > 0) {
                TRY(lr_inner(h, X, j0, ld, P, np, ld, nullptr, C));
                lr_panel_update_kernel<<<nb, THREADS, 0, h->stream>>>(X, m, ld, j0, np, j0, C);
                h->launches += 1;
            }
            lr_panel_gram_kernel<<<nblk, THREADS, 0, h->stream>>>(X, m, ld, j0, np, gpart);
            lr_panel_chol_kernel<<<1, LR_CHOL_THREADS, 0, h->stream>>>(gpart, nblk, np, n0, pass == 0 ? 1 : 0, scale2, T);
            lr_panel_apply_kernel<<<nb, THREADS, 0, h->stream>>>(X, m, ld, j0, np, T);
            h->launches += 3;
        }
    }
    KCHECK();
    return CPD_OK;
}
}  // namespace

extern "C" int cpd_nonrigid_begin(cpd_ctx* h, double beta, double lmd, double sigma2, double w) {
    if (!h) return fail(CPD_ERR_ARG, "null handle");
    if (!h->have_source || !h->have_target) return fail(CPD_ERR_STATE, "source and target must both be set");
    if (!(beta > 0.0) || !(sigma2 > 0.0) || !(w >= 0.0 && w < 1.0)) return fail(CPD_ERR_ARG, "bad beta/sigma2/w");
    CU(cudaSetDevice(h->device));
    TRY(load_cusolver());
    h->nr_ready = false;
    const long long m = h->m;
    TRY(nonrigid_common_begin(h, lmd, sigma2, w));
    if (!h->d_G) TRY(dev_alloc(&h->d_G, (size_t)m * m));
    if (!h->d_A) TRY(dev_alloc(&h->d_A, (size_t)m * m));
    TRY(solver_workspace(h, m, h->d_A));
    TRY(ensure_stats(h));
    const DevState& hs = h->h_state;
    dim3 grid((unsigned)m, blocks_for(m));
    nr_gram_kernel<<<grid, THREADS, 0, h->stream>>>(h->d_yc, hs.cy[0], hs.cy[1], hs.cy[2], m, h->dim, (float)(2.0 * beta),
                                                  h->d_G);
    KCHECK();
    h->launches += 1;
    h->lr_rank = 0;
    h->nr_ready = true;
    return CPD_OK;
}

// NonRigidCPD with G ~= Q Bc Q^T of rank `rank` (lowrank.cuh): randomised range finder with `power_iters` subspace
// iterations on products G X formed by the pair kernel; nothing of size M x M is ever stored.
extern "C" int cpd_nonrigid_lowrank_begin(cpd_ctx* h, double beta, double lmd, double sigma2, double w, int rank, int power_iters,
                                          uint64_t seed) {
    if (!h) return fail(CPD_ERR_ARG, "null handle");
    if (!h->have_source || !h->have_target) return fail(CPD_ERR_STATE, "source and target must both be set");
    if (!(beta > 0.0) || !(sigma2 > 0.0) || !(w >= 0.0 && w < 1.0)) return fail(CPD_ERR_ARG, "bad beta/sigma2/w");
    if (rank < 1 || rank > LR_MAX_RANK) return fail(CPD_ERR_ARG, "rank must be in 1..%d, got %d", LR_MAX_RANK, rank);
    if (power_iters < 0 || power_iters > 8) return fail(CPD_ERR_ARG, "power_iters must be in 0..8, got %d", power_iters);
    CU(cudaSetDevice(h->device));
    TRY(load_cusolver());
    h->nr_ready = false;
    const long long m = h->m, ld = h->mpad;
    if (rank > m) rank = (int)m;
    TRY(nonrigid_common_begin(h, lmd, sigma2, w));
    if (h->lr_m != m || h->lr_cap < rank) {
        TRY(dev_alloc(&h->d_lr_pts, (size_t)ld));
        TRY(dev_alloc(&h->d_lr_Q, (size_t)rank * ld));
        TRY(dev_alloc(&h->d_lr_X, (size_t)rank * ld));
        TRY(dev_alloc(&h->d_lr_coef, (size_t)3 * LR_SLICES * (rank + 1)));
        h->lr_part_cap = std::max<size_t>(std::max<size_t>((size_t)LR_SLICES * rank * rank, (size_t)4 << 20),      // >= 32 MB of slice partials
                                          (size_t)((m + LR_NARROW_PTS - 1) / LR_NARROW_PTS) * rank * 4);
        TRY(dev_alloc(&h->d_lr_part, h->lr_part_cap));
        TRY(dev_alloc(&h->d_lr_panel, (size_t)LR_MAX_RANK * LR_PANEL + 2 * LR_PANEL + LR_PANEL * LR_PANEL +
                                          (size_t)LR_GRAM_BLOCKS * LR_PANEL * LR_PANEL));
        TRY(dev_alloc(&h->d_lr_Bc, (size_t)rank * rank));
        TRY(dev_alloc(&h->d_lr_Lt, (size_t)rank * rank));
        TRY(dev_alloc(&h->d_lr_S, (size_t)rank * rank));
        TRY(dev_alloc(&h->d_lr_R, (size_t)rank * 3));
        TRY(dev_alloc(&h->d_lr_sys, (size_t)rank * rank));
        TRY(dev_alloc(&h->d_lr_rhs, (size_t)rank * 3));
        TRY(dev_alloc(&h->d_lr_c, 1));
        h->lr_m = m;
        h->lr_cap = rank;
    }
    TRY(solver_workspace(h, rank, h->d_lr_sys));
    CU(cudaFuncSetAttribute(lr_spd_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lr_spd_smem_bytes(LR_SPD_MAX_RANK)));
    TRY(ensure_stats(h));
    const DevState& hs = h->h_state;
    lr_pack_kernel<<<blocks_for(ld), THREADS, 0, h->stream>>>(h->d_yc, hs.cy[0], hs.cy[1], hs.cy[2], m, ld,
                                                              (float)sqrt(LOG2E / (2.0 * beta)), h->d_lr_pts);
    dim3 rgrid(blocks_for(m), (unsigned)rank);
    lr_random_kernel<<<rgrid, THREADS, 0, h->stream>>>(h->d_lr_X, m, ld, rank, (unsigned long long)seed);
    KCHECK();
    h->launches += 2;
    // Q <- orth(G Omega); then `power_iters` times Q <- orth(G Q); finally X = G Q and Bc = Q^T X (symmetrised).
    // With profiling on (cpd_set_profiling) the three phases are timed with events: cpd_lowrank_setup_times.
    std::vector<cudaEvent_t> ev;
    std::vector<int> phase;                  // phase of the interval that ENDS at the matching event: 0 products, 1 orth, 2 core
    auto tick = [&](int ph) {
        if (!h->profiling) return;
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, h->stream);
        ev.push_back(e);
        phase.push_back(ph);
    };
    tick(-1);
    TRY(lr_gram_apply(h, h->d_lr_X, h->d_lr_Q, rank));
    tick(0);
    TRY(lr_orthonormalise(h, h->d_lr_Q, rank));
    tick(1);
    for (int it = 0; it < power_iters; ++it) {
        TRY(lr_gram_apply(h, h->d_lr_Q, h->d_lr_X, rank));
        tick(0);
        std::swap(h->d_lr_Q, h->d_lr_X);
        TRY(lr_orthonormalise(h, h->d_lr_Q, rank));
        tick(1);
    }
    TRY(lr_gram_apply(h, h->d_lr_Q, h->d_lr_X, rank));
    tick(0);
    TRY(lr_inner_sym(h, h->d_lr_Q, rank, ld, h->d_lr_X, ld, nullptr, h->d_lr_Bc));
    { const char* e = getenv("CPD_B200_LR_CORE"); h->lr_spd = rank <= LR_SPD_MAX_RANK && !(e && !strcmp(e, "lu")); }
    if (h->lr_spd) TRY(lr_spd_form(h, rank));
    tick(2);
    if (!ev.empty()) {
        h->lr_setup_ms[0] = h->lr_setup_ms[1] = h->lr_setup_ms[2] = 0.0f;
        cudaEventSynchronize(ev.back());
        for (size_t i = 1; i < ev.size(); ++i) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, ev[i - 1], ev[i]) == cudaSuccess) h->lr_setup_ms[phase[i]] += ms;
        }
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
    }
    h->lr_rank = rank;
    h->nr_ready = true;
    return CPD_OK;
}

// Start another registration with the SAME source (one template, many targets): W = 0, T = Y, new sigma2 / w / lmd, priors off --
// G (or its low-rank factors) stays.  The caller guarantees that the source set on this handle is the one the last
// cpd_nonrigid_*begin saw (cpd_set_source with identical coordinates is fine: the internal order is deterministic).
extern "C" int cpd_nonrigid_restart(cpd_ctx* h, double lmd, double sigma2, double w) {
    if (!h) return fail(CPD_ERR_ARG, "null handle");
    if (!h->nr_ready) return fail(CPD_ERR_STATE, "cpd_nonrigid_begin has not been called");
    if (!h->have_source || !h->have_target) return fail(CPD_ERR_STATE, "source and target must both be set");
    if (!(sigma2 > 0.0) || !(w >= 0.0 && w < 1.0)) return fail(CPD_ERR_ARG, "bad sigma2/w");
    if (h->nr_m != h->m || (h->lr_rank > 0 && h->lr_m != h->m)) return fail(CPD_ERR_STATE, "the source size changed since cpd_nonrigid_begin");
    CU(cudaSetDevice(h->device));
    return nonrigid_common_begin(h, lmd, sigma2, w);
}

// Correspondence priors of ConstrainedNonRigidCPD (cpd.py:364-374, 390-396): p1_tilde (m) and px_tilde (m x D) in the caller's
// order, alpha > 0.  Both NULL: priors off.  Valid until the next cpd_nonrigid_*begin with another source size.
extern "C" int cpd_nonrigid_set_prior(cpd_ctx* h, double alpha, const double* p1_tilde, const double* px_tilde) {
    if (!h) return fail(CPD_ERR_ARG, "null handle");
    if (!h->nr_ready) return fail(CPD_ERR_STATE, "cpd_nonrigid_begin has not been called");
    CU(cudaSetDevice(h->device));
    if (!p1_tilde && !px_tilde) { h->prior_on = false; return CPD_OK; }
    if (!p1_tilde || !px_tilde) return fail(CPD_ERR_ARG, "p1_tilde and px_tilde must be given together");
    if (!(alpha > 0.0)) return fail(CPD_ERR_ARG, "alpha must be positive, got %g", alpha);
    const long long m = h->m;
    if (!h->d_p1t || h->prior_m != m) {
        TRY(dev_alloc(&h->d_p1t, (size_t)m));
        TRY(dev_alloc(&h->d_pxt, (size_t)m * 3));
        h->prior_m = m;
    }
    // caller's order -> internal (Morton) order
    CU(cudaMemcpyAsync(h->d_outM, p1_tilde, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    gather1_kernel<<<blocks_for(m), THREADS, 0, h->stream>>>(h->d_outM, h->d_perm_src, m, h->d_p1t);
    if (h->raw_cap < (size_t)m * 3) { TRY(dev_alloc(&h->d_raw, (size_t)m * 3)); h->raw_cap = (size_t)m * 3; }
    TRY(upload_cloud(h, px_tilde, m, h->d_raw));
    gather3_kernel<<<blocks_for(m), THREADS, 0, h->stream>>>(h->d_raw, h->d_perm_src, m, 0.0, 0.0, 0.0, h->d_pxt);
    KCHECK();
    CU(cudaStreamSynchronize(h->stream));
    h->launches += 2;
    h->prior_alpha = alpha;
    h->prior_on = true;
    return CPD_OK;
}

namespace {
// From d_p1 / d_pxc (and the priors) to W and the moved source d_ts2 = Y + G W: the linear system of cpd.py:296, dense LU or
// the K x K form of lowrank.cuh.  sigma2 of the PREVIOUS iteration is read from the device state.
int nonrigid_solve(cpd_ctx* h) {
    const long long m = h->m;
    TRY(ensure_stats(h));
    const DevState& hs = h->h_state;
    const int nbs = (int)blocks_for(m);
    // weights and right-hand side of cpd.py:296 (with priors: cpd.py:390-396)
    const double* wgt = h->d_p1;
    if (h->prior_on) {
        nr_weight_kernel<<<nbs, THREADS, 0, h->stream>>>(h->d_p1, h->d_p1t, &h->d_state->sigma2, h->prior_alpha, m, h->d_wgt);
        h->launches += 1;
        wgt = h->d_wgt;
    }
    nr_rhs_kernel<<<nbs, THREADS, 0, h->stream>>>(h->d_state, h->d_p1, h->d_pxc, h->d_yc, h->prior_on ? h->d_p1t : nullptr,
                                                  h->prior_on ? h->d_pxt : nullptr, h->prior_alpha, m, h->d_B);
    KCHECK();
    h->launches += 1;
    if (h->lr_rank == 0) {
        dim3 grid((unsigned)m, blocks_for(m));
        nr_system_kernel<<<grid, THREADS, 0, h->stream>>>(h->d_G, wgt, &h->d_state->sigma2, h->nr_lmd, m, h->d_A);
        KCHECK();
        SOLV(g_sol.Xgetrf(h->sol, h->sol_params, m, m, CUDA_R_64F_, h->d_A, m, h->d_ipiv, CUDA_R_64F_, h->d_work, h->work_dev, h->h_work,
                          h->work_host, h->d_info));
        SOLV(g_sol.Xgetrs(h->sol, h->sol_params, CUBLAS_OP_T_, m, 3, CUDA_R_64F_, h->d_A, m, h->d_ipiv, CUDA_R_64F_, h->d_B, m, h->d_info));
        nr_unpack_kernel<<<nbs, THREADS, 0, h->stream>>>(h->d_B, m, h->d_W);
        nr_apply_kernel<<<(unsigned)((m + 7) / 8), THREADS, 0, h->stream>>>(h->d_G, h->d_W, h->d_yc, hs.cy[0], hs.cy[1], hs.cy[2], m,
                                                                            h->d_ts2);
        h->launches += 3;
    } else {
        const int k = h->lr_rank;
        const long long ld = h->mpad;
        const double* Qf = h->lr_spd ? h->d_lr_X : h->d_lr_Q;          // the factor of the iteration: Qt = Q L (lr_spd_form) or Q
        TRY(lr_inner_sym(h, Qf, k, ld, Qf, ld, wgt, h->d_lr_S));                 // S = Q^T diag(wgt) Q
        {   // R = Q^T F (F is [3][m]): the narrow product kernel, block partials merged in a fixed order
            const int nblk = (int)((m + LR_NARROW_PTS - 1) / LR_NARROW_PTS);
            if ((size_t)nblk * k * 3 > h->lr_part_cap) return fail(CPD_ERR_STATE, "partial buffer too small for Q^T F");
            lr_inner_narrow_kernel<<<dim3((unsigned)((k + 7) / 8), (unsigned)nblk), THREADS, 0, h->stream>>>(Qf, k, ld, h->d_B, 3, m, m,
                                                                                                         h->d_lr_part);
            lr_merge_kernel<<<blocks_for((long long)k * 3 * 8), THREADS, 0, h->stream>>>(h->d_lr_part, nblk, k, 3, h->d_lr_R);
            KCHECK();
            h->launches += 2;
        }
        if (h->lr_spd) {
            // (c I + St) Z = Rt: symmetric positive definite, one CTA (lr_spd_solve_kernel)
            lr_spd_solve_kernel<<<1, LR_SPD_THREADS, lr_spd_smem_bytes(k), h->stream>>>(h->d_lr_S, h->d_lr_R, k, &h->d_state->sigma2, h->nr_lmd,
                                                                                       h->d_lr_rhs, h->d_lr_c);
            KCHECK();
        } else {
            lr_system_kernel<<<blocks_for((long long)k * k + 3 * k), THREADS, 0, h->stream>>>(h->d_lr_Bc, h->d_lr_S, h->d_lr_R, k,
                                                                                              &h->d_state->sigma2, h->nr_lmd, h->d_lr_sys,
                                                                                              h->d_lr_rhs, h->d_lr_c);
            KCHECK();
            SOLV(g_sol.Xgetrf(h->sol, h->sol_params, k, k, CUDA_R_64F_, h->d_lr_sys, k, h->d_ipiv, CUDA_R_64F_, h->d_work, h->work_dev,
                              h->h_work, h->work_host, h->d_info));
            SOLV(g_sol.Xgetrs(h->sol, h->sol_params, CUBLAS_OP_T_, k, 3, CUDA_R_64F_, h->d_lr_sys, k, h->d_ipiv, CUDA_R_64F_, h->d_lr_rhs, k,
                              h->d_info));
        }
        lr_apply_kernel<<<nbs, THREADS, 0, h->stream>>>(Qf, m, ld, k, h->d_lr_rhs, h->d_yc, hs.cy[0], hs.cy[1], hs.cy[2], h->d_ts2);
        h->lr_w_stale = true;          // W = (F - diag(wgt) Q Z) / c is formed when somebody asks for it (cpd_nonrigid_get)
        h->launches += 2;
    }
    return CPD_OK;
}
}  // namespace

// one EM iteration of probreg/cpd.py:111-113 for NonRigidCPD: E-step on T = Y + G W, solve cpd.py:296, sigma2 cpd.py:298-301
extern "C" int cpd_nonrigid_step(cpd_ctx* h, double* sigma2_out) {
    if (!h) return fail(CPD_ERR_ARG, "null handle");
    if (!h->nr_ready) return fail(CPD_ERR_STATE, "cpd_nonrigid_begin has not been called");
    CU(cudaSetDevice(h->device));
    const long long m = h->m;
    TRY(launch_estep(h, &h->d_state->sigma2, &h->d_state->w, h->d_ts));
    const int nbs = (int)blocks_for(m), nbt = (int)blocks_for(h->npad);
    moments_kernel<0><<<1, 256, 0, h->stream>>>(h->d_state, h->d_mom_src, nbs, RM_SRC, h->d_mom_tgt, nbt, RM_TGT, h->d_mom);
    h->launches += 1;
    if (h->comm) {   // every rank solves the same (global) system
        TRY(allreduce(h, h->d_p1, (size_t)m));
        TRY(allreduce(h, h->d_pxc, (size_t)m * 3));
        TRY(allreduce(h, h->d_mom, MOM_PAD));
    }
    TRY(nonrigid_solve(h));
    nr_resid_kernel<<<nbs, THREADS, 0, h->stream>>>(h->d_state, h->d_p1, h->d_pxc, h->d_ts, h->d_ts2, m, h->d_nrpart);
    nr_sigma_kernel<<<1, 32, 0, h->stream>>>(h->d_state, h->d_nrpart, nbs, h->d_mom);
    mark(h, 6);          // cpd_stage_times: [5] is the whole M-step (moments, system, LU / K x K solve, T, sigma2)
    KCHECK();
    h->launches += 2;
    std::swap(h->d_ts, h->d_ts2);
    if (sigma2_out) {
        CU(cudaMemcpyAsync(h->h_pin + 56, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        cpd_params p;
        TRY(read_params(h, &p));
        const int info = *reinterpret_cast<const int*>(h->h_pin + 56);
        if (info != 0) return fail(CPD_ERR_STATE, "LU factorisation of the non-rigid system failed (info = %d)", info);
        *sigma2_out = p.sigma2;
    }
    return CPD_OK;
}

// NonRigidCPD._maximization_step / ConstrainedNonRigidCPD._maximization_step (cpd.py:284-303 / 376-404) from a caller-supplied
// EstepResult: pt1 (n_local), p1 (m), px (m x D) as returned by cpd_estep, sigma2_p the variance the E-step was run with.
// Needs cpd_nonrigid_begin / cpd_nonrigid_lowrank_begin (and optionally cpd_nonrigid_set_prior) on this handle.  The new W
// and moved source are read with cpd_nonrigid_get.  sigma2 by the reference's three traces, FP64.
extern "C" int cpd_nonrigid_mstep(cpd_ctx* h, const double* pt1, const double* p1, const double* px, double sigma2_p, double* sigma2_out) {
    if (!h || !pt1 || !p1 || !px) return fail(CPD_ERR_ARG, "null argument");
    if (!h->nr_ready) return fail(CPD_ERR_STATE, "cpd_nonrigid_begin has not been called");
    if (!(sigma2_p > 0.0)) return fail(CPD_ERR_ARG, "sigma2_p must be positive, got %g", sigma2_p);
    CU(cudaSetDevice(h->device));
    TRY(prepare(h));
    const long long m = h->m, n = h->n;
    // caller's order -> internal (Morton) order, px -> centred px~ (as cpd_mstep does)
    CU(cudaMemcpyAsync(h->d_outN, pt1, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    gather1_kernel<<<blocks_for(n), THREADS, 0, h->stream>>>(h->d_outN, h->d_perm_tgt, n, h->d_pt1);
    CU(cudaMemcpyAsync(h->d_outM, p1, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    gather1_kernel<<<blocks_for(m), THREADS, 0, h->stream>>>(h->d_outM, h->d_perm_src, m, h->d_p1);
    if (h->raw_cap < (size_t)m * 3) { TRY(dev_alloc(&h->d_raw, (size_t)m * 3)); h->raw_cap = (size_t)m * 3; }
    TRY(upload_cloud(h, px, m, h->d_raw));
    gather3_kernel<<<blocks_for(m), THREADS, 0, h->stream>>>(h->d_raw, h->d_perm_src, m, 0.0, 0.0, 0.0, h->d_px);
    centre_px_kernel<<<blocks_for(m), THREADS, 0, h->stream>>>(h->d_state, h->d_p1, h->d_px, (int)m, h->d_pxc);
    h->h_pin[36] = sigma2_p;
    CU(cudaMemcpyAsync(&h->d_state->sigma2, h->h_pin + 36, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    KCHECK();
    h->launches += 4;
    TRY(nonrigid_solve(h));
    const unsigned nb = blocks_for(std::max(m, n));
    if (h->sums_cap < (size_t)nb * 4 + 4) { TRY(dev_alloc(&h->d_sums, (size_t)nb * 4 + 4)); h->sums_cap = (size_t)nb * 4 + 4; }
    nr_traces_kernel<<<nb, THREADS, 0, h->stream>>>(h->d_state, h->d_pt1, h->d_xc, n, h->d_p1, h->d_pxc, h->d_ts2, m, h->d_sums + 4);
    reduce_cols_kernel<<<1, 4 * 32, 0, h->stream>>>(h->d_sums + 4, (int)nb, 4, h->d_sums);
    KCHECK();
    if (h->comm) TRY(allreduce(h, h->d_sums, 1));            // pt1 / x are per shard; p1, px, T are global already
    nr_sigma_api_kernel<<<1, 32, 0, h->stream>>>(h->d_state, h->d_sums);
    KCHECK();
    h->launches += 3;
    std::swap(h->d_ts, h->d_ts2);
    CU(cudaMemcpyAsync(h->h_pin + 56, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    cpd_params p;
    TRY(read_params(h, &p));
    const int info = *reinterpret_cast<const int*>(h->h_pin + 56);
    if (info != 0) return fail(CPD_ERR_STATE, "LU factorisation of the non-rigid system failed (info = %d)", info);
    if (sigma2_out) *sigma2_out = p.sigma2;
    return CPD_OK;
}

// W (m x D) of the current NonRigidTransformation, and optionally the moved source T = Y + G W (m x D)
extern "C" int cpd_nonrigid_get(cpd_ctx* h, double* w_out, double* moved_out) {
    if (!h) return fail(CPD_ERR_ARG, "null handle");
    if (!h->nr_ready) return fail(CPD_ERR_STATE, "cpd_nonrigid_begin has not been called");
    CU(cudaSetDevice(h->device));
    if (w_out) {
        if (h->lr_rank > 0 && h->lr_w_stale) {      // d_B still holds F, d_ts the moved source and d_lr_c the c of the last solve
            TRY(ensure_stats(h));
            const DevState& hs = h->h_state;
            lr_w_kernel<<<blocks_for(h->m), THREADS, 0, h->stream>>>(h->d_B, h->prior_on ? h->d_wgt : h->d_p1, h->d_ts, h->d_yc, hs.cy[0],
                                                                     hs.cy[1], hs.cy[2], h->m, h->d_lr_c, h->d_W);
            h->lr_w_stale = false;
            h->launches += 1;
        }
        scatter_kernel<<<blocks_for(h->m), THREADS, 0, h->stream>>>(h->d_W, h->d_perm_src, h->m, 3, h->d_outM);
        TRY(download_cloud(h, h->d_outM, h->m, w_out));
        CU(cudaStreamSynchronize(h->stream));
        h->launches += 1;
    }
    if (moved_out) {
        scatter_kernel<<<blocks_for(h->m), THREADS, 0, h->stream>>>(h->d_ts, h->d_perm_src, h->m, 3, h->d_outM);
        TRY(download_cloud(h, h->d_outM, h->m, moved_out));
        CU(cudaStreamSynchronize(h->stream));
        h->launches += 1;
    }
    KCHECK();
    return CPD_OK;
}

// Duration of the last cpd_nonrigid_lowrank_begin that ran with profiling on, by phase: [0] the G X products, [1] the
// orthonormalisations, [2] Bc = Q^T (G Q).
extern "C" int cpd_lowrank_setup_times(cpd_ctx* h, float ms[3]) {
    if (!h || !ms) return fail(CPD_ERR_ARG, "null argument");
    for (int k = 0; k < 3; ++k) ms[k] = h->lr_setup_ms[k];
    return CPD_OK;
}

// The factors of the low-rank path in the caller's point order: q_out (m x rank, row-major, orthonormal columns) and
// bcore_out (rank x rank, symmetric) with G ~= q bcore q^T; rank_out receives the rank in use (<= the one asked for).
extern "C" int cpd_nonrigid_lowrank_get(cpd_ctx* h, int* rank_out, double* q_out, double* bcore_out) {
    if (!h) return fail(CPD_ERR_ARG, "null handle");
    if (!h->nr_ready || h->lr_rank == 0) return fail(CPD_ERR_STATE, "cpd_nonrigid_lowrank_begin has not been called");
    CU(cudaSetDevice(h->device));
    const int k = h->lr_rank;
    if (rank_out) *rank_out = k;
    if (q_out) {
        const size_t need = (size_t)h->m * k;
        if (h->lr_out_cap < need) { TRY(dev_alloc(&h->d_lr_out, need)); h->lr_out_cap = need; }
        lr_export_kernel<<<blocks_for(h->m), THREADS, 0, h->stream>>>(h->d_lr_Q, h->d_perm_src, h->m, h->mpad, k, h->d_lr_out);
        KCHECK();
        h->launches += 1;
        CU(cudaMemcpyAsync(q_out, h->d_lr_out, need * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    if (bcore_out) CU(cudaMemcpyAsync(bcore_out, h->d_lr_Bc, (size_t)k * k * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return CPD_OK;
}

