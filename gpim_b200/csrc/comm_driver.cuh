// comm_driver.cuh -- gpg_comm_* / gpg_bcast_factor / gpg_allgather_pred / gpg_predict_sharded / gpg_acq_sweep_sharded:
// the multi-GPU entry points of the C ABI.  Included at the end of gpgrid.cu: uses its predict and
// acquisition drivers.  See comm.cuh for the run-time NCCL binding and the stream discipline.

extern "C" int gpg_comm_unique_id(void *id_host) {
    GPG_REQUIRE(id_host != nullptr, "id_host is NULL");
    comm::Api *a = comm::api();
    if (!a) { gpg_set_error("libnccl.so.2 could not be loaded: %s", dlerror()); return GPG_ECUDA; }
    comm::UniqueId id;
    GPG_NCCL_CHECK(a->GetUniqueId(&id));
    memcpy(id_host, &id, sizeof(id));
    return GPG_OK;
}

extern "C" int gpg_comm_destroy(gpg_handle_t h) {
    if (!h || !h->comm) return GPG_OK;
    DeviceGuard device_guard(h->device);
    comm::State *st = reinterpret_cast<comm::State *>(h->comm);
    if (st->stream) cudaStreamSynchronize(st->stream);
    if (st->comm && comm::api()) comm::api()->CommDestroy(st->comm);
    for (cudaEvent_t e : st->ev_chunk) cudaEventDestroy(e);
    for (cudaEvent_t e : {st->ev_in, st->ev_out, st->ev_small}) if (e) cudaEventDestroy(e);
    if (st->stream) cudaStreamDestroy(st->stream);
    delete st;
    h->comm = nullptr;
    return GPG_OK;
}

extern "C" int gpg_comm_init(gpg_handle_t h, int nranks, int rank, const void *id_host) {
    GPG_REQUIRE(h && id_host, "NULL argument");
    GPG_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "rank not in [0, nranks)");
    comm::Api *a = comm::api();
    if (!a) { gpg_set_error("libnccl.so.2 could not be loaded: %s", dlerror()); return GPG_ECUDA; }
    DeviceGuard device_guard(h->device);
    GPG_TRY(gpg_comm_destroy(h));
    comm::State *st = new comm::State();
    h->comm = st;
    st->nranks = nranks; st->rank = rank;
    comm::UniqueId id;
    memcpy(&id, id_host, sizeof(id));
    GPG_CUDA_CHECK(cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking));
    for (cudaEvent_t *e : {&st->ev_in, &st->ev_out, &st->ev_small}) GPG_CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    GPG_NCCL_CHECK(a->CommInitRank(&st->comm, nranks, id, rank));
    return GPG_OK;
}

extern "C" int gpg_comm_info(gpg_handle_t h, int *nranks_out, int *rank_out) {
    GPG_REQUIRE(h != nullptr, "handle is NULL");
    comm::State *st = reinterpret_cast<comm::State *>(h->comm);
    if (nranks_out) *nranks_out = (st && st->comm) ? st->nranks : 0;
    if (rank_out) *rank_out = (st && st->comm) ? st->rank : -1;
    return GPG_OK;
}

extern "C" int gpg_predict_uses_planes(gpg_handle_t h, int dtype, int64_t N, int have_planes) {
    return h && predict_wants_tc(h, dtype, N, have_planes != 0) ? 1 : 0;
}

static size_t dtype_bytes(int dtype) { return dtype == GPG_F64 ? 8 : 4; }

// everything but the N x N part of the cache, as ONE fused NCCL launch on the communication stream
static int bcast_small(comm::State *st, int dtype, int d, int64_t N, void *theta, void *X, void *alpha, float *scales,
                       int32_t *info, int root) {
    const size_t eb = dtype_bytes(dtype);
    GPG_NCCL_CHECK(comm::api()->GroupStart());
    int rc = comm::bcast_bytes(st, theta, (3 + d) * eb, root);
    if (rc == GPG_OK) rc = comm::bcast_bytes(st, X, (size_t)N * d * eb, root);
    if (rc == GPG_OK) rc = comm::bcast_bytes(st, alpha, (size_t)N * eb, root);
    if (rc == GPG_OK) rc = comm::bcast_bytes(st, scales, SC_COUNT * sizeof(float), root);
    if (rc == GPG_OK) rc = comm::bcast_bytes(st, info, sizeof(int32_t), root);
    GPG_NCCL_CHECK(comm::api()->GroupEnd());
    return rc;
}

// rows [r0, r1) of the N x N part: both fp16 planes (tcgen05 route) or the Linv rows themselves; one fused launch
static int bcast_rows(comm::State *st, int dtype, int64_t N, int64_t ld, void *Linv, void *wsplit, bool planes, int64_t r0,
                      int64_t r1, int root) {
    if (r1 <= r0) return GPG_OK;
    int rc = GPG_OK;
    GPG_NCCL_CHECK(comm::api()->GroupStart());
    if (planes) {
        __half *hi = (__half *)wsplit, *lo = hi + (size_t)N * ld;
        rc = comm::bcast_bytes(st, hi + r0 * ld, (size_t)(r1 - r0) * ld * 2, root);
        if (rc == GPG_OK) rc = comm::bcast_bytes(st, lo + r0 * ld, (size_t)(r1 - r0) * ld * 2, root);
    } else {
        rc = comm::bcast_bytes(st, (unsigned char *)Linv + (size_t)r0 * ld * dtype_bytes(dtype), (size_t)(r1 - r0) * ld * dtype_bytes(dtype), root);
    }
    GPG_NCCL_CHECK(comm::api()->GroupEnd());
    return rc;
}

extern "C" int gpg_bcast_factor(gpg_handle_t h, int dtype, int d, int64_t N, int64_t ld, void *theta, void *X, void *Linv,
                                void *alpha, void *wsplit, float *scales, int32_t *info, int root, void *stream) {
    GPG_REQUIRE(h && theta && X && alpha && info, "NULL argument");
    GPG_REQUIRE(dtype == GPG_F32 || dtype == GPG_F64, "unknown dtype");
    GPG_REQUIRE(N > 0 && ld >= N && d >= 1 && d <= GPG_MAX_D, "bad size");
    comm::State *st;
    GPG_TRY(comm::need(h, &st));
    GPG_REQUIRE(root >= 0 && root < st->nranks, "root not in [0, nranks)");
    DeviceGuard device_guard(h->device);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const bool planes = predict_wants_tc(h, dtype, N, wsplit != nullptr);
    GPG_REQUIRE(planes || Linv != nullptr, "Linv is NULL but the SIMT route will read it");
    GPG_TRY(comm::fork_from(st, s));
    GPG_TRY(bcast_small(st, dtype, d, N, theta, X, alpha, dtype == GPG_F32 ? scales : nullptr, info, root));
    GPG_TRY(bcast_rows(st, dtype, N, ld, Linv, wsplit, planes, 0, N, root));
    return comm::join_into(st, s);
}

extern "C" int gpg_allgather_pred(gpg_handle_t h, int dtype, const void *pred_local, int64_t count, void *pred_all,
                                  void *stream) {
    GPG_REQUIRE(h && pred_local && pred_all, "NULL argument");
    GPG_REQUIRE(dtype == GPG_F32 || dtype == GPG_F64, "unknown dtype");
    GPG_REQUIRE(count >= 0, "negative count");
    comm::State *st;
    GPG_TRY(comm::need(h, &st));
    DeviceGuard device_guard(h->device);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (count == 0) return GPG_OK;
    GPG_TRY(comm::fork_from(st, s));
    GPG_NCCL_CHECK(comm::api()->AllGather(pred_local, pred_all, (size_t)count * dtype_bytes(dtype), comm::NCCL_UINT8, st->comm,
                                          st->stream));
    return comm::join_into(st, s);
}

extern "C" int gpg_predict_sharded(gpg_handle_t h, int dtype, int kernel_id, int d, void *theta, void *X, int64_t N,
                                   void *Linv, int64_t ld, void *alpha, void *wsplit, float *scales, int32_t *info, int root,
                                   const void *Xs_local, int64_t M_local, void *pred_local, int64_t M_pad, void *pred_all,
                                   void *stream) {
    GPG_REQUIRE(h && theta && X && alpha && info && pred_local, "NULL argument");
    GPG_REQUIRE(dtype == GPG_F32 || dtype == GPG_F64, "unknown dtype");
    GPG_REQUIRE(N > 0 && ld >= N && d >= 1 && d <= GPG_MAX_D, "bad size");
    GPG_REQUIRE(M_local >= 0 && M_pad >= M_local && (M_local == 0 || Xs_local != nullptr), "bad tile");
    comm::State *st;
    GPG_TRY(comm::need(h, &st));
    GPG_REQUIRE(root >= 0 && root < st->nranks, "root not in [0, nranks)");
    DeviceGuard device_guard(h->device);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const bool planes = predict_wants_tc(h, dtype, N, wsplit != nullptr);
    GPG_REQUIRE(planes || Linv != nullptr, "Linv is NULL but the SIMT route will read it");
    const size_t eb = dtype_bytes(dtype);
    void *mean = pred_local, *sd = (unsigned char *)pred_local + (size_t)M_pad * eb;

    // ---- broadcast on the communication stream, pipelined by row blocks when the tcgen05 route will consume it
    GPG_TRY(comm::fork_from(st, s));
    GPG_TRY(bcast_small(st, dtype, d, N, theta, X, alpha, dtype == GPG_F32 ? scales : nullptr, info, root));
    GPG_CUDA_CHECK(cudaEventRecord(st->ev_small, st->stream));
    int nchunks = 1;
    if (planes && st->nranks > 1) {             // ~ 32 MB per block, at most 16 blocks, boundaries on n-block edges
        const int64_t rows_per = std::max<int64_t>(tc::BN, gpg_align_up((size_t)(((int64_t)32 << 20) / (ld * 4) + 1), tc::BN));
        nchunks = (int)std::min<int64_t>(16, (N + rows_per - 1) / rows_per);
    }
    std::vector<int64_t> row_end(nchunks);
    {
        const int64_t nblocks = (N + tc::BN - 1) / tc::BN;
        // Linv is lower triangular: the bytes that matter grow with the row index, but whole rows travel; equal row counts
        for (int c = 0; c < nchunks; ++c) row_end[c] = std::min<int64_t>(N, ((nblocks * (c + 1)) / nchunks) * tc::BN);
        row_end[nchunks - 1] = N;
    }
    while ((int)st->ev_chunk.size() < nchunks) {
        cudaEvent_t e;
        GPG_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        st->ev_chunk.push_back(e);
    }
    for (int c = 0; c < nchunks; ++c) {
        GPG_TRY(bcast_rows(st, dtype, N, ld, Linv, wsplit, planes, c == 0 ? 0 : row_end[c - 1], row_end[c], root));
        GPG_CUDA_CHECK(cudaEventRecord(st->ev_chunk[c], st->stream));
    }

    // ---- this rank's tile on the caller's stream
    GPG_CUDA_CHECK(cudaStreamWaitEvent(s, st->ev_small, 0));
    PlaneArrival arrival;
    arrival.nchunks = nchunks; arrival.row_end = row_end.data(); arrival.ev = st->ev_chunk.data();
    const bool pipelined = planes && nchunks > 1 && M_local > 0;
    if (!pipelined)
        for (int c = 0; c < nchunks; ++c) GPG_CUDA_CHECK(cudaStreamWaitEvent(s, st->ev_chunk[c], 0));
    if (M_local > 0) {
        if (dtype == GPG_F32)
            GPG_TRY(predict_entry<float>(h, kernel_id, d, (const float *)theta, (const float *)X, N, (const float *)Linv, ld,
                                         (const float *)alpha, planes ? wsplit : nullptr, scales, (const float *)Xs_local, nullptr,
                                         nullptr, 0, M_local, (float *)mean, (float *)sd, s, pipelined ? &arrival : nullptr));
        else
            GPG_TRY(predict_entry<double>(h, kernel_id, d, (const double *)theta, (const double *)X, N, (const double *)Linv, ld,
                                          (const double *)alpha, nullptr, nullptr, (const double *)Xs_local, nullptr, nullptr, 0,
                                          M_local, (double *)mean, (double *)sd, s));
    }
    // ---- one all-gather of the {mean, sd} tiles
    if (pred_all && M_pad > 0) {
        GPG_TRY(comm::fork_from(st, s));
        GPG_NCCL_CHECK(comm::api()->AllGather(pred_local, pred_all, 2 * (size_t)M_pad * eb, comm::NCCL_UINT8, st->comm, st->stream));
    }
    return comm::join_into(st, s);
}

// K6 sharded: local sweep + top-k on this rank's tile, a k * nranks-element gather, the merge on every rank.
template <typename T>
static int acq_sharded_entry(gpg_handle_s *h, comm::State *st, int acq_id, const T *mean, const T *sd, const T *mask,
                             int64_t M_local, int64_t idx_offset, double mu_best, double xi, double alpha, double beta, int k,
                             T *topk_val, int64_t *topk_idx, int32_t *count, T *acq_out, cudaStream_t s) {
    constexpr int CH = 2048;
    StageTimer stt(h, GPG_ST_ACQ, s);
    const size_t total = (size_t)k * st->nranks;
    Cand<T> *best, *extra;
    GPG_TRY(acq_local_topk<T>(h, acq_id, mean, sd, mask, M_local, idx_offset, mu_best, xi, alpha, beta, k,
                              acq_out, total, &best, &extra, s));
    Cand<T> *gathered = extra, *merged = extra + total + CH;
    GPG_TRY(comm::fork_from(st, s));
    GPG_NCCL_CHECK(comm::api()->AllGather(best, gathered, (size_t)k * sizeof(Cand<T>), comm::NCCL_UINT8, st->comm, st->stream));
    GPG_TRY(comm::join_into(st, s));
    const Cand<T> *in = gathered;
    int64_t n = (int64_t)total;
    Cand<T> *out = merged;
    while (true) {                                   // k * nranks <= 1024 * nranks: one or two rounds
        const int64_t nblk = (n + CH - 1) / CH;
        topk_round_kernel<T><<<(unsigned)nblk, 1024, CH * sizeof(Cand<T>), s>>>(in, n, k, out, nullptr, nullptr);
        GPG_LAUNCH_CHECK(h);
        if (nblk == 1) break;
        in = out;
        n = nblk * k;
        out = (out == merged) ? gathered : merged;
    }
    topk_emit_kernel<T><<<1, 256, 0, s>>>(out, k, topk_val, topk_idx, count);
    GPG_LAUNCH_CHECK(h);
    return GPG_OK;
}

extern "C" int gpg_acq_sweep_sharded(gpg_handle_t h, int dtype, int acq_id, const void *mean_local, const void *sd_local,
                                     const void *mask_local, int64_t M_local, int64_t idx_offset, double mu_best, double xi,
                                     double alpha, double beta, int k, void *topk_val, int64_t *topk_idx, int32_t *count_out,
                                     void *acq_out_local, void *stream) {
    GPG_REQUIRE(h && topk_val && topk_idx && count_out, "NULL argument");
    GPG_REQUIRE(M_local >= 0 && (M_local == 0 || (mean_local && sd_local)), "bad tile");
    GPG_REQUIRE(k >= 1 && k <= 1024, "k must be in 1..1024");
    GPG_REQUIRE(acq_id >= 0 && acq_id <= 2, "unknown acquisition id");
    comm::State *st;
    GPG_TRY(comm::need(h, &st));
    DeviceGuard device_guard(h->device);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == GPG_F32)
        return acq_sharded_entry<float>(h, st, acq_id, (const float *)mean_local, (const float *)sd_local, (const float *)mask_local,
                                        M_local, idx_offset, mu_best, xi, alpha, beta, k, (float *)topk_val, topk_idx, count_out,
                                        (float *)acq_out_local, s);
    if (dtype == GPG_F64)
        return acq_sharded_entry<double>(h, st, acq_id, (const double *)mean_local, (const double *)sd_local,
                                         (const double *)mask_local, M_local, idx_offset, mu_best, xi, alpha, beta, k,
                                         (double *)topk_val, topk_idx, count_out, (double *)acq_out_local, s);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}
