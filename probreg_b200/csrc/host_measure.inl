// host_measure.inl: measurement helpers of bench.py (events, stage times, L2 flush, issue-rate probes) -- part of the single translation unit cpd_b200.cu (included at its end; uses its handle type, error
// macros and helpers).  Split out for readability only.
extern "C" int cpd_timer_start(cpd_ctx* h) {
    if (!h) return fail(CPD_ERR_ARG, "null handle");
    CU(cudaEventRecord(h->ev0, h->stream));
    return CPD_OK;
}
extern "C" int cpd_timer_stop(cpd_ctx* h, float* ms) {
    if (!h || !ms) return fail(CPD_ERR_ARG, "null argument");
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaEventSynchronize(h->ev1));
    CU(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return CPD_OK;
}
extern "C" int cpd_sync(cpd_ctx* h) {
    if (!h) return fail(CPD_ERR_ARG, "null handle");
    CU(cudaStreamSynchronize(h->stream));
    if (h->d_p2p) {          // a fused P2P exchange may have given up on a peer since the last read of the state
        CU(cudaMemcpyAsync(h->h_pin + 40, &h->d_state->err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (*reinterpret_cast<const int*>(h->h_pin + 40))
            return fail(CPD_ERR_STATE, "a peer rank did not deliver its moments within the P2P exchange timeout");
    }
    return CPD_OK;
}
extern "C" int cpd_event_record(cpd_ctx* h, int idx) {
    if (!h || idx < 0 || idx >= 8192) return fail(CPD_ERR_ARG, "bad event slot %d", idx);
    if ((int)h->pool.size() <= idx) h->pool.resize((size_t)idx + 1, nullptr);
    if (!h->pool[idx]) CU(cudaEventCreate(&h->pool[idx]));
    CU(cudaEventRecord(h->pool[idx], h->stream));
    return CPD_OK;
}
extern "C" int cpd_event_elapsed(cpd_ctx* h, int a, int b, float* ms) {
    if (!h || !ms || a < 0 || b < 0 || a >= (int)h->pool.size() || b >= (int)h->pool.size() || !h->pool[a] || !h->pool[b])
        return fail(CPD_ERR_ARG, "event slot not recorded");
    CU(cudaEventSynchronize(h->pool[b]));
    CU(cudaEventElapsedTime(ms, h->pool[a], h->pool[b]));
    return CPD_OK;
}
extern "C" int cpd_set_profiling(cpd_ctx* h, int on) {
    if (!h) return fail(CPD_ERR_ARG, "null handle");
    h->profiling = on != 0;
    return CPD_OK;
}
extern "C" int cpd_stage_times(cpd_ctx* h, float ms[6]) {
    if (!h || !ms) return fail(CPD_ERR_ARG, "null argument");
    CU(cudaEventSynchronize(h->sev[6]));
    for (int k = 0; k < 6; ++k) CU(cudaEventElapsedTime(&ms[k], h->sev[k], h->sev[k + 1]));
    return CPD_OK;
}
extern "C" int64_t cpd_launch_count(cpd_ctx* h) { return h ? h->launches : 0; }

extern "C" int cpd_flush_l2(cpd_ctx* h, int64_t bytes) {
    if (!h) return fail(CPD_ERR_ARG, "null handle");
    const size_t want = bytes > 0 ? (size_t)bytes : ((size_t)256 << 20);
    if (h->flush_cap < want) {
        if (h->d_flush) cudaFree(h->d_flush);
        h->d_flush = nullptr;
        CU(cudaMalloc(&h->d_flush, want));
        h->flush_cap = want;
    }
    CU(cudaMemsetAsync(h->d_flush, 0x5a, want, h->stream));
    return CPD_OK;
}

extern "C" int cpd_microbench(int device, double out[9]) {
    if (!out) return fail(CPD_ERR_ARG, "null argument");
    if (cpd_device_count() == 0) return fail(CPD_ERR_CUDA, "no CUDA device");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    float* d = nullptr;
    long long* dc = nullptr;
    TRY(dev_alloc(&d, 16));
    TRY(dev_alloc(&dc, 2));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, iters = 100000;
    float ms = 0.f;
#define TIME_PROBE(KERNEL, IT)                       \
    KERNEL<<<blocks, 256>>>(d, 500, 1.0f);           \
    CU(cudaEventRecord(e0));                         \
    KERNEL<<<blocks, 256>>>(d, (IT), 1.0f);          \
    CU(cudaEventRecord(e1));                         \
    CU(cudaEventSynchronize(e1));                    \
    CU(cudaEventElapsedTime(&ms, e0, e1));
    const double threads = (double)blocks * 256.0;
    TIME_PROBE(probe_ffma_kernel, iters);
    out[0] = threads * iters * 16.0 * 2.0 / (ms * 1e-3) / 1e12;              // FFMA TFLOP/s
    TIME_PROBE(probe_mufu_kernel, iters / 4);
    out[1] = threads * (iters / 4) * 8.0 / (ms * 1e-3) / 1e9;                // MUFU.EX2 Gop/s
    TIME_PROBE(probe_ffma2_kernel, iters);
    out[4] = threads * iters * 8.0 * 4.0 / (ms * 1e-3) / 1e12;               // FFMA2 TFLOP/s
    // pairs/s of an (11 FP32 + 1 MUFU) mix, scalar and packed (6 FFMA2 ~ 12 FP32 per 2 pairs), and (7 + 1)
    TIME_PROBE((probe_mix_kernel<11, false>), iters / 8);
    out[5] = threads * (iters / 8) * 8.0 / (ms * 1e-3) / 1e9;                // Gpairs/s, scalar 11+1
    TIME_PROBE((probe_mix_kernel<5, true>), iters / 8);
    out[6] = threads * (iters / 8) * 8.0 / (ms * 1e-3) / 1e9;                // Gpairs/s, packed (5 FFMA2 + FMUL) + 1 MUFU
    TIME_PROBE((probe_mix_kernel<7, false>), iters / 8);
    out[7] = threads * (iters / 8) * 8.0 / (ms * 1e-3) / 1e9;                // Gpairs/s, scalar 7+1
    TIME_PROBE(probe_ffma_mixed_kernel, iters);
    out[8] = threads * iters * 6.0 * (4.0 + 2.0) / (ms * 1e-3) / 1e12;       // TFLOP/s of FFMA2 + FFMA interleaved 1:1
#undef TIME_PROBE
    probe_clock_kernel<<<1, 1>>>(dc);
    long long hc[2];
    CU(cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost));
    out[2] = (double)hc[0] / ((double)hc[1] * 1e-9) / 1e6;
    out[3] = prop.multiProcessorCount;
    cudaFree(d); cudaFree(dc);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return CPD_OK;
}
