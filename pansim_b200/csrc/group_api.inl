// group_api.inl -- several GPUs in ONE process (included inside extern "C" of pansim_b200.cu).
//
// SURVEY.md 8b "Threading": the reference is one process whose Population methods are called
// sequentially from the main thread (main.rs:429-528); a host that keeps that shape drives all GPUs
// of the box through a pansim_group: one shard context per device (column shards of the core
// alignment, accessory matrix replicated), one NCCL communicator per device from ncclCommInitAll.
// Every group call enqueues on all shards first and synchronises afterwards, so the devices work
// concurrently although a single host thread issues everything.
//
// Exchanges: none in the generation step; the distance pass sums the per-pair partial
// core counts -- ncclAllReduce for sampled pairs (1.2 MB at P = 1e5), and for the exact all-pairs
// mode of BASELINE config 5 a ncclReduceScatter per row block on a second stream, overlapped with
// the kernels of the next block (pansim_group_all_pairs), so every device ends up with, and copies
// out, only its 1/n slice of the summed counts.

struct pansim_group {
    std::vector<pansim_ctx *> ctx;
    std::vector<ncclComm_t> comms;
    std::string err;
    // all-pairs pipeline: second stream + events per shard, two sets of pinned host vectors
    std::vector<cudaStream_t> stream_comm;
    std::vector<cudaEvent_t> ev_compute, ev_done[2];
    uint32_t *h_cnt[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    size_t h_cap = 0;
    // device time of the reduce-scatter on shard 0 (timed events around every block's collective) and wall time of the last walk
    cudaEvent_t ev_rs0[2] = {nullptr, nullptr}, ev_rs1[2] = {nullptr, nullptr};
    float last_nccl_ms = 0.f, last_wall_ms = 0.f;
    uint64_t last_pairs = 0;
};

namespace {

#define GFAIL(g, code, ...)                                  \
    do {                                                     \
        char _b[512];                                        \
        snprintf(_b, sizeof _b, __VA_ARGS__);                \
        (g)->err = _b;                                       \
        return (code);                                       \
    } while (0)

// error of shard i -> error of the group
int gfail_from(pansim_group *g, int i, int rc)
{
    g->err = "shard " + std::to_string(i) + ": " + g->ctx[i]->err;
    return rc;
}

int group_sync_all(pansim_group *g, int code, const char *what)
{
    int first = 0;
    for (size_t i = 0; i < g->ctx.size(); i++) {
        pansim_ctx *c = g->ctx[i];
        cudaSetDevice(c->cfg.device);
        int rc = check_device_flag(c, code, what);
        if (rc && !first) first = gfail_from(g, (int)i, rc);
    }
    return first;
}

// sum the partial core counts of all shards in place (every shard ends with the whole sum)
int group_allreduce(pansim_group *g, uint32_t *const *bufs, size_t n)
{
    if (g->ctx.size() < 2 || !n) return 0;
    NcclApi &api = nccl_api();
    ncclResult_t r = api.GroupStart();
    for (size_t i = 0; i < g->ctx.size() && r == ncclSuccess; i++)
        r = api.AllReduce(bufs[i], bufs[i], n, ncclUint32, ncclSum, g->comms[i], g->ctx[i]->ps);
    const ncclResult_t r2 = api.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) GFAIL(g, PANSIM_ERR_CUDA, "ncclAllReduce over the shards failed: %s", api.GetErrorString(r));
    return 0;
}

void group_shards(uint64_t core_size, int n, std::vector<std::pair<uint64_t, uint64_t>> &out)
{
    // whole regions per shard, as even as the alignment allows (pansim_b200/sharding.py: column_shards)
    const uint64_t regions = (core_size + PANSIM_SITE_ALIGN - 1) / PANSIM_SITE_ALIGN;
    const uint64_t base = regions / n, extra = regions % n;
    uint64_t r0 = 0;
    for (int i = 0; i < n; i++) {
        const uint64_t nr = base + ((uint64_t)i < extra ? 1 : 0);
        out.push_back({std::min(core_size, r0 * PANSIM_SITE_ALIGN), std::min(core_size, (r0 + nr) * PANSIM_SITE_ALIGN)});
        r0 += nr;
    }
}

}  // namespace

// the column shard [begin, end) of shard i of n: whole 8192-site regions, as even as the alignment allows
// (no device needed; pansim_b200/sharding.py: column_shards computes the same split)
int pansim_shard_bounds(uint64_t core_size, int n_shards, int shard, uint64_t *site_begin, uint64_t *site_end)
{
    if (n_shards < 1 || shard < 0 || shard >= n_shards || !site_begin || !site_end) return PANSIM_ERR_INVALID;
    std::vector<std::pair<uint64_t, uint64_t>> s;
    group_shards(core_size, n_shards, s);
    *site_begin = s[shard].first;
    *site_end = s[shard].second;
    return 0;
}

int pansim_group_create(const pansim_config *cfg, int n_devices, const int *devices, pansim_group **out)
{
    if (!cfg || !out || n_devices < 1) { g_create_error = "pansim_group_create: bad argument"; return PANSIM_ERR_INVALID; }
    *out = nullptr;
    if (cfg->site_begin || cfg->site_end) { g_create_error = "pansim_group_create: the configuration must describe the whole alignment"; return PANSIM_ERR_INVALID; }
    pansim_group *g = new pansim_group();
    std::vector<std::pair<uint64_t, uint64_t>> shards;
    group_shards(cfg->core_size, n_devices, shards);
    std::vector<int> devs(n_devices);
    for (int i = 0; i < n_devices; i++) devs[i] = devices ? devices[i] : i;
    for (int i = 0; i < n_devices; i++) {
        pansim_config ci = *cfg;
        ci.device = devs[i];
        if (n_devices > 1) {
            if (shards[i].first == shards[i].second) {
                g_create_error = "pansim_group_create: more devices than 8192-site regions";
                pansim_group_destroy(g);
                return PANSIM_ERR_INVALID;
            }
            ci.site_begin = shards[i].first;
            ci.site_end = shards[i].second;
        }
        pansim_ctx *c = nullptr;
        const int rc = pansim_create(&ci, &c);
        if (rc) { pansim_group_destroy(g); return rc; }       // g_create_error holds the message
        g->ctx.push_back(c);
    }
    if (n_devices > 1) {
        NcclApi &api = nccl_api();
        if (!api.ok()) { g_create_error = api.error; pansim_group_destroy(g); return PANSIM_ERR_CUDA; }
        g->comms.assign(n_devices, nullptr);
        const ncclResult_t r = api.CommInitAll(g->comms.data(), n_devices, devs.data());
        if (r != ncclSuccess) {
            g_create_error = std::string("ncclCommInitAll failed: ") + api.GetErrorString(r);
            g->comms.clear();
            pansim_group_destroy(g);
            return PANSIM_ERR_CUDA;
        }
        for (int i = 0; i < n_devices; i++) {
            g->ctx[i]->comm = g->comms[i];
            g->ctx[i]->comm_size = n_devices;
            g->ctx[i]->comm_rank = i;
            g->ctx[i]->comm_owned = false;
        }
    }
    *out = g;
    return PANSIM_OK;
}

void pansim_group_destroy(pansim_group *g)
{
    if (!g) return;
    for (size_t i = 0; i < g->ctx.size(); i++) {
        pansim_ctx *c = g->ctx[i];
        cudaSetDevice(c->cfg.device);
        cudaDeviceSynchronize();
        if (i < g->stream_comm.size() && g->stream_comm[i]) cudaStreamDestroy(g->stream_comm[i]);
        if (i < g->ev_compute.size() && g->ev_compute[i]) cudaEventDestroy(g->ev_compute[i]);
        if (i == 0) for (int s = 0; s < 2; s++) { if (g->ev_rs0[s]) cudaEventDestroy(g->ev_rs0[s]); if (g->ev_rs1[s]) cudaEventDestroy(g->ev_rs1[s]); }
        for (int s = 0; s < 2; s++)
            if (i < g->ev_done[s].size() && g->ev_done[s][i]) cudaEventDestroy(g->ev_done[s][i]);
    }
    for (size_t i = 0; i < g->comms.size(); i++)
        if (g->comms[i]) nccl_api().CommDestroy(g->comms[i]);
    for (pansim_ctx *c : g->ctx) {
        c->comm = nullptr;
        pansim_destroy(c);
    }
    for (int s = 0; s < 2; s++)
        for (int k = 0; k < 3; k++)
            if (g->h_cnt[s][k]) cudaFreeHost(g->h_cnt[s][k]);
    delete g;
}

const char *pansim_group_last_error(const pansim_group *g) { return g ? g->err.c_str() : g_create_error.c_str(); }

int pansim_group_size(const pansim_group *g) { return g ? (int)g->ctx.size() : 0; }

pansim_ctx *pansim_group_ctx(pansim_group *g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }

int pansim_group_set_initial(pansim_group *g, const uint8_t *core_row_onehot, const uint8_t *acc_row)
{
    if (!g) return PANSIM_ERR_INVALID;
    for (size_t i = 0; i < g->ctx.size(); i++)
        if (int rc = pansim_set_initial(g->ctx[i], core_row_onehot, acc_row)) return gfail_from(g, (int)i, rc);
    return 0;
}

int pansim_group_set_selection(pansim_group *g, const double *s)
{
    if (!g) return PANSIM_ERR_INVALID;
    for (size_t i = 0; i < g->ctx.size(); i++)
        if (int rc = pansim_set_selection(g->ctx[i], s)) return gfail_from(g, (int)i, rc);
    return 0;
}

// main.rs:435-464 for n generations on every shard: parents, flips and HGT are recomputed identically
// from the same counters on every device, so nothing is exchanged
int pansim_group_run_generations(pansim_group *g, uint32_t gen0, uint32_t n)
{
    if (!g) return PANSIM_ERR_INVALID;
    for (size_t i = 0; i < g->ctx.size(); i++)
        if (int rc = run_generations_enqueue(g->ctx[i], gen0, n)) return gfail_from(g, (int)i, rc);
    return group_sync_all(g, PANSIM_ERR_WEIGHTS, PANSIM_WEIGHTS_MSG);
}

// Population::pairwise_distances over the whole alignment (population.rs:787-837): partial core counts
// per shard, ncclAllReduce, read-back from shard 0 (the accessory counts are complete on every shard)
int pansim_group_pair_counts(pansim_group *g, const uint32_t *r1, const uint32_t *r2, size_t P, uint32_t *core_diff,
                             uint32_t *inter, uint32_t *uni)
{
    if (!g || (P && (!r1 || !r2))) return PANSIM_ERR_INVALID;
    if (P == 0) return 0;
    std::vector<uint32_t *> bufs(g->ctx.size());
    for (size_t i = 0; i < g->ctx.size(); i++) {
        pansim_ctx *c = g->ctx[i];
        int rc = require_state(c);
        if (!rc) rc = cudaSetDevice(c->cfg.device) == cudaSuccess ? 0 : PANSIM_ERR_CUDA;
        if (!rc) rc = ensure_pairs(c, P);
        if (!rc) rc = pair_counts_impl(c, r1, r2, P, core_diff ? c->d_cd : nullptr, (i == 0 && (inter || uni)) ? c->d_in : nullptr,
                                       (i == 0 && (inter || uni)) ? c->d_un : nullptr);
        if (rc) return gfail_from(g, (int)i, rc);
        bufs[i] = c->d_cd;
    }
    if (core_diff)
        if (int rc = group_allreduce(g, bufs.data(), P)) return rc;
    pansim_ctx *c0 = g->ctx[0];
    cudaSetDevice(c0->cfg.device);
    if (core_diff) cudaMemcpyAsync(core_diff, c0->d_cd, P * 4, cudaMemcpyDeviceToHost, c0->stream);
    if (inter) cudaMemcpyAsync(inter, c0->d_in, P * 4, cudaMemcpyDeviceToHost, c0->stream);
    if (uni) cudaMemcpyAsync(uni, c0->d_un, P * 4, cudaMemcpyDeviceToHost, c0->stream);
    return group_sync_all(g, PANSIM_ERR_WEIGHTS, PANSIM_WEIGHTS_MSG);
}

// the --print_dist loop (main.rs:429-519) on all shards: per generation the step, the distance pass,
// the all-reduce of the core counts and the statistics kernel (shard 0); stats_out[n][4]
int pansim_group_run_generations_stats(pansim_group *g, uint32_t gen0, uint32_t n, const uint32_t *r1, const uint32_t *r2,
                                       size_t P, double *stats_out)
{
    if (!g || !stats_out || !P || !r1 || !r2) return PANSIM_ERR_INVALID;
    if (n == 0) return 0;
    const size_t ns = g->ctx.size();
    for (size_t i = 0; i < ns; i++) {
        pansim_ctx *c = g->ctx[i];
        int rc = require_state(c);
        if (!rc) rc = cudaSetDevice(c->cfg.device) == cudaSuccess ? 0 : PANSIM_ERR_CUDA;
        if (!rc) rc = ensure_pairs(c, P);
        if (!rc) rc = ensure_stats(c, n, P, true);
        if (!rc) rc = pair_prepare(c, r1, r2, P);
        if (rc) return gfail_from(g, (int)i, rc);
        timing_begin(c);
    }
    std::vector<uint32_t *> bufs(ns);
    std::vector<uint32_t *> in_(ns), un_(ns);
    for (uint32_t gen = 0; gen < n; gen++) {
        const int slot = (int)(gen & 1u);
        for (size_t i = 0; i < ns; i++) {
            pansim_ctx *c = g->ctx[i];
            cudaSetDevice(c->cfg.device);
            if (int rc = stats_generation(c, gen0 + gen, slot, P, i == 0, &bufs[i], &in_[i], &un_[i])) return gfail_from(g, (int)i, rc);
        }
        if (int rc = group_allreduce(g, bufs.data(), P)) return rc;
        for (size_t i = 0; i < ns; i++) {
            pansim_ctx *c = g->ctx[i];
            cudaSetDevice(c->cfg.device);
            if (i == 0) {
                if (int rc = launch_pair_stats(c, slot, P, bufs[0], in_[0], un_[0], c->d_stats + (size_t)gen * 4)) return gfail_from(g, 0, rc);
            } else {
                cudaEventRecord(c->ev_pairs[slot], c->ps);          // pass done on this shard (launch_pair_stats records it on shard 0)
                c->ev_pairs_valid[slot] = true;
            }
        }
    }
    for (size_t i = 0; i < ns; i++) {
        pansim_ctx *c = g->ctx[i];
        cudaSetDevice(c->cfg.device);
        if (int rc = stats_batch_end(c)) return gfail_from(g, (int)i, rc);
        timing_end(c);
    }
    pansim_ctx *c0 = g->ctx[0];
    cudaSetDevice(c0->cfg.device);
    for (int i = 0; i < 2; i++)
        if (c0->ev_stats_valid[i]) cudaStreamWaitEvent(c0->stream, c0->ev_stats[i], 0);
    cudaMemcpyAsync(c0->h_stats, c0->d_stats, (size_t)n * 4 * sizeof(double), cudaMemcpyDeviceToHost, c0->stream);
    if (int rc = group_sync_all(g, PANSIM_ERR_WEIGHTS, PANSIM_WEIGHTS_MSG)) return rc;
    memcpy(stats_out, c0->h_stats, (size_t)n * 4 * sizeof(double));
    return 0;
}

// ---- exact all-pairs mode over the group (BASELINE config 5) ---------------------------------
// Row blocks of about chunk_pairs pairs, each (i, j) with i < j exactly once, ordered by i then j.
// Per block: every shard computes the partial core counts of ALL pairs of the block over its
// columns; a ncclReduceScatter on the shard's second stream leaves shard r with the summed counts
// of slice r, which it copies into the pinned host vector; the accessory counts of the block come
// from one shard (they rotate). While that traffic is in flight the kernels of the next block run.
// `cb` is called once per block, in order, with host vectors that stay valid during the call.
int pansim_group_all_pairs(pansim_group *g, size_t chunk_pairs, pansim_pairs_cb cb, void *user)
{
    if (!g || !cb) return PANSIM_ERR_INVALID;
    const size_t ns = g->ctx.size();
    const uint32_t N = g->ctx[0]->N;
    if (chunk_pairs < N) chunk_pairs = N;
    if (chunk_pairs > 0x3FFFFFFFull) chunk_pairs = 0x3FFFFFFFull;
    NcclApi &api = nccl_api();
    // row blocks
    struct Block { uint32_t i0, i1; size_t P; };
    std::vector<Block> blocks;
    for (uint32_t i0 = 0; i0 + 1 < N;) {
        uint32_t i1 = i0;
        size_t np = 0;
        while (i1 + 1 < N && (np == 0 || np + (N - 1 - i1) <= chunk_pairs)) { np += N - 1 - i1; i1++; }
        blocks.push_back({i0, i1, np});
        i0 = i1;
    }
    if (blocks.empty()) return 0;
    size_t maxP = 0;
    for (const Block &b : blocks) maxP = std::max(maxP, b.P);
    const size_t slice_max = (maxP + ns - 1) / ns, padded_max = slice_max * ns;
    // resources
    if (g->stream_comm.empty()) {
        g->stream_comm.assign(ns, nullptr);
        g->ev_compute.assign(ns, nullptr);
        g->ev_done[0].assign(ns, nullptr);
        g->ev_done[1].assign(ns, nullptr);
        for (size_t i = 0; i < ns; i++) {
            cudaSetDevice(g->ctx[i]->cfg.device);
            if (cudaStreamCreateWithFlags(&g->stream_comm[i], cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&g->ev_compute[i], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&g->ev_done[0][i], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&g->ev_done[1][i], cudaEventDisableTiming) != cudaSuccess)
                GFAIL(g, PANSIM_ERR_CUDA, "stream / event creation for the all-pairs pipeline failed");
        }
    }
    if (padded_max > g->h_cap) {
        for (int s = 0; s < 2; s++)
            for (int k = 0; k < 3; k++) {
                if (g->h_cnt[s][k]) cudaFreeHost(g->h_cnt[s][k]);
                g->h_cnt[s][k] = nullptr;
                if (cudaMallocHost(&g->h_cnt[s][k], padded_max * 4) != cudaSuccess) GFAIL(g, PANSIM_ERR_NOMEM, "pinned host vectors for %zu pairs failed", padded_max);
            }
        g->h_cap = padded_max;
    }
    for (size_t i = 0; i < ns; i++) {
        pansim_ctx *c = g->ctx[i];
        int rc = require_state(c);
        if (!rc) rc = cudaSetDevice(c->cfg.device) == cudaSuccess ? 0 : PANSIM_ERR_CUDA;
        if (!rc) rc = ensure_pairs(c, padded_max);
        if (!rc) rc = ensure_stats(c, 1, padded_max, true);       // second set of count vectors
        if (rc) return gfail_from(g, (int)i, rc);
    }
    if (!g->ev_rs0[0]) {
        cudaSetDevice(g->ctx[0]->cfg.device);
        for (int s = 0; s < 2; s++) { cudaEventCreate(&g->ev_rs0[s]); cudaEventCreate(&g->ev_rs1[s]); }
    }
    g->last_nccl_ms = 0.f; g->last_pairs = 0;
    const auto wall0 = std::chrono::steady_clock::now();
    auto deliver = [&](size_t b) -> int {
        const int set = (int)(b & 1u);
        for (size_t i = 0; i < ns; i++) {
            cudaSetDevice(g->ctx[i]->cfg.device);
            if (cudaEventSynchronize(g->ev_done[set][i]) != cudaSuccess) GFAIL(g, PANSIM_ERR_CUDA, "all-pairs block %zu failed on shard %zu: %s", b, i, cudaGetErrorString(cudaGetLastError()));
        }
        if (ns > 1) {
            float ms = 0.f;
            cudaSetDevice(g->ctx[0]->cfg.device);
            if (cudaEventElapsedTime(&ms, g->ev_rs0[set], g->ev_rs1[set]) == cudaSuccess) g->last_nccl_ms += ms;
        }
        g->last_pairs += blocks[b].P;
        return cb(user, blocks[b].i0, blocks[b].i1, blocks[b].P, g->h_cnt[set][0], g->h_cnt[set][1], g->h_cnt[set][2]);
    };
    for (size_t b = 0; b < blocks.size(); b++) {
        const int set = (int)(b & 1u);
        const Block &blk = blocks[b];
        const size_t slice = (blk.P + ns - 1) / ns;
        const size_t acc_shard = b % ns;
        for (size_t i = 0; i < ns; i++) {
            pansim_ctx *c = g->ctx[i];
            cudaSetDevice(c->cfg.device);
            uint32_t *cd = set ? c->d_cnt2[0] : c->d_cd, *in = set ? c->d_cnt2[1] : c->d_in, *un = set ? c->d_cnt2[2] : c->d_un;
            // the transfers of block b-2 out of this set of vectors have been delivered (deliver(b-2) waited for them)
            size_t P = 0;
            int rc = pair_counts_rows_impl(c, blk.i0, blk.i1, cd, i == acc_shard ? in : nullptr, i == acc_shard ? un : nullptr, &P);
            if (rc) return gfail_from(g, (int)i, rc);
            cudaEventRecord(g->ev_compute[i], c->stream);
            cudaStreamWaitEvent(g->stream_comm[i], g->ev_compute[i], 0);
        }
        if (ns > 1) {
            cudaSetDevice(g->ctx[0]->cfg.device);
            cudaEventRecord(g->ev_rs0[set], g->stream_comm[0]);
            ncclResult_t r = api.GroupStart();
            for (size_t i = 0; i < ns && r == ncclSuccess; i++) {
                pansim_ctx *c = g->ctx[i];
                uint32_t *cd = set ? c->d_cnt2[0] : c->d_cd;
                r = api.ReduceScatter(cd, cd + i * slice, slice, ncclUint32, ncclSum, g->comms[i], g->stream_comm[i]);
            }
            const ncclResult_t r2 = api.GroupEnd();
            cudaSetDevice(g->ctx[0]->cfg.device);
            cudaEventRecord(g->ev_rs1[set], g->stream_comm[0]);
            if (r == ncclSuccess) r = r2;
            if (r != ncclSuccess) GFAIL(g, PANSIM_ERR_CUDA, "ncclReduceScatter of the all-pairs counts failed: %s", api.GetErrorString(r));
        }
        for (size_t i = 0; i < ns; i++) {
            pansim_ctx *c = g->ctx[i];
            cudaSetDevice(c->cfg.device);
            uint32_t *cd = set ? c->d_cnt2[0] : c->d_cd, *in = set ? c->d_cnt2[1] : c->d_in, *un = set ? c->d_cnt2[2] : c->d_un;
            const size_t lo = std::min(blk.P, i * slice), hi = std::min(blk.P, (i + 1) * slice);
            if (hi > lo) cudaMemcpyAsync(g->h_cnt[set][0] + lo, cd + lo, (hi - lo) * 4, cudaMemcpyDeviceToHost, g->stream_comm[i]);
            if (i == acc_shard) {
                cudaMemcpyAsync(g->h_cnt[set][1], in, blk.P * 4, cudaMemcpyDeviceToHost, g->stream_comm[i]);
                cudaMemcpyAsync(g->h_cnt[set][2], un, blk.P * 4, cudaMemcpyDeviceToHost, g->stream_comm[i]);
            }
            // block b+2 re-uses this set of device and host vectors: it is enqueued after deliver(b), which
            // waits for this event on the host
            cudaEventRecord(g->ev_done[set][i], g->stream_comm[i]);
        }
        if (b > 0)
            if (int rc = deliver(b - 1)) return rc;
    }
    if (int rc = deliver(blocks.size() - 1)) return rc;
    const int rc_end = group_sync_all(g, PANSIM_ERR_WEIGHTS, PANSIM_WEIGHTS_MSG);
    g->last_wall_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    return rc_end;
}

// wall time of the last pansim_group_all_pairs walk (callbacks included), device time shard 0 spent in the
// reduce-scatters (they run on the second stream, overlapped with the kernels of the next block), pairs walked
int pansim_group_all_pairs_timing(pansim_group *g, float *wall_ms, float *nccl_ms, uint64_t *pairs)
{
    if (!g) return PANSIM_ERR_INVALID;
    if (wall_ms) *wall_ms = g->last_wall_ms;
    if (nccl_ms) *nccl_ms = g->last_nccl_ms;
    if (pairs) *pairs = g->last_pairs;
    return 0;
}

// ---- replicated / assembled state ------------------------------------------------------------
int pansim_group_gene_counts(pansim_group *g, uint32_t *counts)
{
    if (!g) return PANSIM_ERR_INVALID;
    if (int rc = pansim_gene_counts(g->ctx[0], counts)) return gfail_from(g, 0, rc);
    return 0;
}

int pansim_group_download_acc(pansim_group *g, uint8_t *acc_out)
{
    if (!g) return PANSIM_ERR_INVALID;
    if (int rc = pansim_download_acc(g->ctx[0], acc_out)) return gfail_from(g, 0, rc);
    return 0;
}

// [N x core_size] one-hot bytes assembled from the column shards
int pansim_group_download_core(pansim_group *g, uint8_t *core_out)
{
    if (!g || !core_out) return PANSIM_ERR_INVALID;
    const uint64_t L = g->ctx[0]->L;
    const uint32_t N = g->ctx[0]->N;
    for (size_t i = 0; i < g->ctx.size(); i++) {
        pansim_ctx *c = g->ctx[i];
        if (c->Ll == L) return pansim_download_core(c, core_out) ? gfail_from(g, (int)i, PANSIM_ERR_CUDA) : 0;
        std::vector<uint8_t> part((size_t)N * c->Ll);
        if (int rc = pansim_download_core(c, part.data())) return gfail_from(g, (int)i, rc);
        for (uint32_t r = 0; r < N; r++) memcpy(core_out + (size_t)r * L + c->site_begin, part.data() + (size_t)r * c->Ll, c->Ll);
    }
    return 0;
}

// rows [row_begin, row_end) of _core_genome.csv (population.rs:877-879): 2 * core_size bytes per row
int pansim_group_export_core_csv(pansim_group *g, uint32_t row_begin, uint32_t row_end, char *out)
{
    if (!g || !out || row_begin > row_end) return PANSIM_ERR_INVALID;
    const uint64_t L = g->ctx[0]->L;
    const uint32_t nr = row_end - row_begin;
    for (size_t i = 0; i < g->ctx.size(); i++) {
        pansim_ctx *c = g->ctx[i];
        if (c->Ll == L) return pansim_export_core_csv(c, row_begin, row_end, out) ? gfail_from(g, (int)i, PANSIM_ERR_CUDA) : 0;
        std::vector<char> part((size_t)nr * 2 * c->Ll);
        if (int rc = pansim_export_core_csv(c, row_begin, row_end, part.data())) return gfail_from(g, (int)i, rc);
        for (uint32_t r = 0; r < nr; r++) {
            char *dst = out + (size_t)r * 2 * L + 2 * c->site_begin;
            memcpy(dst, part.data() + (size_t)r * 2 * c->Ll, 2 * c->Ll);
            if (c->site_end != L) dst[2 * c->Ll - 1] = ',';      // a shard's row ends with a newline; inside a whole row it is a separator
        }
    }
    return 0;
}
