// host_multi.inl: multi-GPU plumbing (NCCL communicator, NVLink peer-memory mailboxes) -- part of the single translation unit cpd_b200.cu (included at its end; uses its handle type, error
// macros and helpers).  Split out for readability only.
extern "C" int cpd_comm_unique_id(char id[128]) {
    TRY(load_nccl());
    nccl_uid u;
    NC(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return CPD_OK;
}

extern "C" int cpd_comm_create(void** comm, int device, int world_size, int rank, const char id[128]) {
    if (!comm || !id) return fail(CPD_ERR_ARG, "null argument");
    if (world_size < 1 || rank < 0 || rank >= world_size) return fail(CPD_ERR_ARG, "bad world_size/rank %d/%d", world_size, rank);
    TRY(load_nccl());
    CU(cudaSetDevice(device));
    nccl_uid u;
    memcpy(u.internal, id, 128);
    nccl_comm c = nullptr;
    NC(g_nccl.CommInitRank(&c, world_size, u, rank));
    *comm = c;
    return CPD_OK;
}

extern "C" int cpd_comm_destroy(void* comm) {
    if (!comm) return CPD_OK;
    TRY(load_nccl());
    NC(g_nccl.CommDestroy((nccl_comm)comm));
    return CPD_OK;
}

extern "C" int cpd_comm_attach(cpd_ctx* h, void* comm, int world_size, int rank) {
    if (!h) return fail(CPD_ERR_ARG, "null handle");
    if (comm && (world_size < 1 || rank < 0 || rank >= world_size)) return fail(CPD_ERR_ARG, "bad world_size/rank %d/%d", world_size, rank);
    h->comm = (nccl_comm)comm;
    h->world = comm ? world_size : 1;
    h->rank = comm ? rank : 0;
    return CPD_OK;
}

extern "C" int cpd_p2p_local_handle(cpd_ctx* h, char out[64]) {
    if (!h || !out) return fail(CPD_ERR_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
    CU(cudaSetDevice(h->device));
    if (!h->d_box) {
        TRY(dev_alloc(&h->d_box, 1));
        CU(cudaMemset(h->d_box, 0, sizeof(P2PMailbox)));
        CU(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t mh;
    CU(cudaIpcGetMemHandle(&mh, h->d_box));
    memcpy(out, &mh, 64);
    return CPD_OK;
}

extern "C" int cpd_p2p_attach(cpd_ctx* h, const char* handles, int world_size, int rank) {
    if (!h || !handles) return fail(CPD_ERR_ARG, "null argument");
    if (world_size < 2 || world_size > P2P_MAX || rank < 0 || rank >= world_size)
        return fail(CPD_ERR_ARG, "bad world_size/rank %d/%d (2..%d ranks)", world_size, rank, P2P_MAX);
    if (!h->d_box) return fail(CPD_ERR_STATE, "cpd_p2p_local_handle must be called first");
    if (h->d_p2p) return fail(CPD_ERR_STATE, "P2P exchange already attached to this handle");
    CU(cudaSetDevice(h->device));
    P2PInfo info;
    memset(&info, 0, sizeof(info));
    info.world = world_size;
    info.rank = rank;
    // A rank waits this long for its peers' moments before it gives up (then DevState.err is set, the M-step is skipped and the
    // next read of the state -- cpd_em_step(out), cpd_em_run, cpd_sync -- returns CPD_ERR_STATE).  Ranks running uneven host work
    // between iterations (a slow callback on one rank) need either a bound above that or CPD_B200_NO_P2P=1 (ncclAllReduce).
    double timeout_s = 120.0;
    if (const char* e = getenv("CPD_B200_P2P_TIMEOUT_S")) { const double v = atof(e); if (v > 0.0) timeout_s = v; }
    info.timeout_ns = (unsigned long long)(timeout_s * 1e9);
    for (int r = 0; r < world_size; ++r) {
        if (r == rank) { info.box[r] = h->d_box; continue; }
        cudaIpcMemHandle_t mh;
        memcpy(&mh, handles + (size_t)r * 64, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&h->peer_ptr[r], mh, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            h->peer_ptr[r] = nullptr;
            cudaGetLastError();
            for (int q = 0; q < r; ++q)
                if (h->peer_ptr[q]) { cudaIpcCloseMemHandle(h->peer_ptr[q]); h->peer_ptr[q] = nullptr; }
            return fail(CPD_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
        }
        info.box[r] = (P2PMailbox*)h->peer_ptr[r];
    }
    TRY(dev_alloc(&h->d_p2p, 1));
    CU(cudaMemcpy(h->d_p2p, &info, sizeof(info), cudaMemcpyHostToDevice));
    h->world = world_size;
    h->rank = rank;
    return CPD_OK;
}

extern "C" int cpd_p2p_detach(cpd_ctx* h) {
    if (!h) return fail(CPD_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    for (int r = 0; r < P2P_MAX; ++r)
        if (h->peer_ptr[r]) { cudaIpcCloseMemHandle(h->peer_ptr[r]); h->peer_ptr[r] = nullptr; }
    if (h->d_p2p) { cudaFree(h->d_p2p); h->d_p2p = nullptr; }
    return CPD_OK;
}

