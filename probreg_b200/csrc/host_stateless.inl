// host_stateless.inl: entry points that need no handle (kernel matrices, direct Gauss transform, squared_kernel_sum) -- part of the single translation unit cpd_b200.cu (included at its end; uses its handle type, error
// macros and helpers).  Split out for readability only.
namespace {
// float32 nx x ny kernel matrix of two host clouds; kind 0: rbf (param = beta), 1: inverse multiquadric (param = c)
int pair_matrix(int kind, int device, const double* x, int64_t nx, const double* y, int64_t ny, int dim, double param, float* out) {
    if (!x || !y || !out || nx < 1 || ny < 1 || dim < 1 || dim > 16) return fail(CPD_ERR_ARG, "bad argument");
    if (cpd_device_count() == 0) return fail(CPD_ERR_CUDA, "no CUDA device: this library has no CPU path");
    CU(cudaSetDevice(device));
    std::vector<float> xf((size_t)nx * dim), yf((size_t)ny * dim);   // the pybind11/Eigen cast to float32 (cc/types.h:19)
    for (size_t i = 0; i < xf.size(); ++i) xf[i] = (float)x[i];
    for (size_t i = 0; i < yf.size(); ++i) yf[i] = (float)y[i];
    DevBuf<float> dx, dy, dout;
    TRY(dx.alloc(xf.size()));
    TRY(dy.alloc(yf.size()));
    TRY(dout.alloc((size_t)nx * ny));
    CU(cudaMemcpy(dx.p, xf.data(), xf.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dy.p, yf.data(), yf.size() * sizeof(float), cudaMemcpyHostToDevice));
    dim3 grid((unsigned)nx, blocks_for(ny));
    if (kind == 0) rbf_kernel_kernel<<<grid, THREADS>>>(dx.p, nx, dy.p, ny, dim, (float)(2.0 * param), dout.p);
    else imq_kernel_kernel<<<grid, THREADS>>>(dx.p, nx, dy.p, ny, dim, (float)param, dout.p);
    KCHECK();
    CU(cudaMemcpy(out, dout.p, (size_t)nx * ny * sizeof(float), cudaMemcpyDeviceToHost));
    return CPD_OK;
}
}  // namespace

extern "C" int cpd_rbf_kernel(int device, const double* x, int64_t nx, const double* y, int64_t ny, int dim, double beta, float* out) {
    return pair_matrix(0, device, x, nx, y, ny, dim, beta, out);
}
extern "C" int cpd_imq_kernel(int device, const double* x, int64_t nx, const double* y, int64_t ny, int dim, double c, float* out) {
    return pair_matrix(1, device, x, nx, y, ny, dim, c, out);
}

// Direct Gauss transform on host arrays (stateless).  weights: k x m row-major, out: k x n.
extern "C" int cpd_gauss_transform(int device, const double* source, int64_t m, const double* target, int64_t n, int dim, double h,
                                   const double* weights, int k, double* out) {
    if (!source || !target || !weights || !out) return fail(CPD_ERR_ARG, "null argument");
    if (m < 1 || n < 1 || k < 1 || dim < 1 || dim > 3 || !(h > 0.0)) return fail(CPD_ERR_ARG, "bad m/n/k/dim/h");
    if (cpd_device_count() == 0) return fail(CPD_ERR_CUDA, "no CUDA device: this library has no CPU path");
    CU(cudaSetDevice(device));
    const int64_t mpad = (m + GT_TILE - 1) / GT_TILE * GT_TILE;
    double c[3] = {0.0, 0.0, 0.0};
    for (int64_t j = 0; j < m; ++j) for (int a = 0; a < dim; ++a) c[a] += source[j * dim + a];
    for (int a = 0; a < dim; ++a) c[a] /= (double)m;
    const double sk = sqrt(LOG2E) / h;                   // exp(-d^2/h^2) = 2^-(sk d)^2
    std::vector<float4> hs((size_t)mpad), ht((size_t)n);
    std::vector<float> hw((size_t)k * mpad, 0.0f);
    for (int64_t j = 0; j < mpad; ++j) {
        float v[3] = {FAR_COORD, FAR_COORD, FAR_COORD};
        if (j < m) for (int a = 0; a < 3; ++a) v[a] = a < dim ? (float)(sk * (source[j * dim + a] - c[a])) : 0.0f;
        hs[(size_t)j] = make_float4(v[0], v[1], v[2], 0.0f);
    }
    for (int64_t i = 0; i < n; ++i) {
        float v[3] = {0.f, 0.f, 0.f};
        for (int a = 0; a < dim; ++a) v[a] = (float)(sk * (target[i * dim + a] - c[a]));
        ht[(size_t)i] = make_float4(v[0], v[1], v[2], 0.0f);
    }
    for (int cc = 0; cc < k; ++cc) for (int64_t j = 0; j < m; ++j) hw[(size_t)cc * mpad + j] = (float)weights[(size_t)cc * m + j];
    DevBuf<float4> ds, dt;
    DevBuf<float> dw;
    DevBuf<double> dout;
    TRY(ds.alloc(hs.size()));
    TRY(dt.alloc(ht.size()));
    TRY(dw.alloc(hw.size()));
    TRY(dout.alloc((size_t)k * n));
    CU(cudaMemcpy(ds.p, hs.data(), hs.size() * sizeof(float4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dt.p, ht.data(), ht.size() * sizeof(float4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dw.p, hw.data(), hw.size() * sizeof(float), cudaMemcpyHostToDevice));
    for (int k0 = 0; k0 < k; k0 += GT_K)
        gauss_transform_kernel<<<blocks_for(n), THREADS>>>(dt.p, (int)n, ds.p, dw.p, (int)mpad, k0, std::min(GT_K, k - k0), dout.p);
    KCHECK();
    CU(cudaMemcpy(out, dout.p, (size_t)k * n * sizeof(double), cudaMemcpyDeviceToHost));
    return CPD_OK;
}

extern "C" int cpd_squared_kernel_sum(int device, const double* x, int64_t nx, const double* y, int64_t ny, int dim, double* out) {
    if (!x || !y || !out) return fail(CPD_ERR_ARG, "null argument");
    cpd_ctx* h = nullptr;
    TRY(cpd_create(&h, device, dim, nullptr));
    int r = cpd_set_source(h, x, nx);
    if (r == CPD_OK) r = cpd_set_target(h, y, ny, ny, nullptr);
    if (r == CPD_OK) r = cpd_sigma2_init(h, out);
    cpd_destroy(h);
    return r;
}

