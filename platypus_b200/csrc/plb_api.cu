// plb_api.cu — host side of the C ABI declared in include/platypus_b200.h.
//
// Owns device memory and launch planning; all arithmetic of the likelihood path runs in the
// kernels of plb_kernels.cuh.  There is deliberately no CPU implementation behind these entry
// points: without a usable CUDA device every call fails with PLB_ERR_CUDA.
#include <cuda_runtime.h>

#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/platypus_b200.h"
#include "plb_kernels.cuh"

using namespace plb;

static thread_local char g_err[512] = "";

static int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

// One OpenMP team size for every host-side parallel region of the library (libgomp rebuilds its team whenever the
// size changes).  Default: the cores this process may run on, divided among the ranks of the node when launched by
// torchrun (LOCAL_WORLD_SIZE), at most 16; PLB_HOST_THREADS overrides.  OMP_NUM_THREADS is deliberately not the cap:
// torchrun exports OMP_NUM_THREADS=1 to every rank, which would leave tile planning on one thread.
static int host_threads() {
    static const int n = [] {
        int v;
        if (getenv("PLB_HOST_THREADS")) {
            v = atoi(getenv("PLB_HOST_THREADS"));
        } else {
            const int cores = omp_get_num_procs();   // honours the affinity mask
            const int ranks = getenv("LOCAL_WORLD_SIZE") ? std::max(1, atoi(getenv("LOCAL_WORLD_SIZE"))) : 1;
            v = std::min(16, cores / ranks);
        }
        return std::max(1, std::min(v, 64));
    }();
    return n;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return set_err(PLB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                           __LINE__);                                                              \
    } while (0)

constexpr int kTimingRing = 32;
constexpr int kMaxChunks = 6;           // pipelined host path: chunks of windows per call
constexpr int kMinChunkWindows = 1024;

struct Block {
    void* p;
    size_t cap;
};

constexpr int kMaxJobs = 2;            // host-path jobs in flight per context (plb_population_submit / _wait)

// Pinned host staging memory of one job slot (tile lists read zero-copy by the kernels, ll offsets, per-round arrays of a
// selection call): valid until the slot's next pin_reset.
struct PinArena {
    std::vector<std::pair<uint8_t*, size_t>> blocks;
    size_t off = 0;
};

// What one in-flight host-path job owns exclusively: its arena, its statistics counters and the events that order its
// copies against its kernels.  Slot 0 also serves the synchronous entry points (device-resident runs, the selection loop).
struct JobSlot {
    PinArena pin;
    Counters* d_ctr = nullptr;
    Counters* h_ctr = nullptr;              // pinned
    cudaEvent_t ev_chunk[kMaxChunks + 1];   // H2D of chunk k done
    cudaEvent_t ev_unpack[kMaxChunks];      // packed input: chunk k's bases are ASCII again
    cudaEvent_t ev_tiles[kMaxChunks];       // chunk k's tile lists are on the device (jobs queued behind another job)
    cudaEvent_t ev_anchor[kMaxChunks];      // chunk k's anchor kernel has finished (software pipeline of the chunks)
    cudaEvent_t ev_kdone[kMaxChunks];       // chunk k's kernels have finished: its outputs may go back
    cudaEvent_t ev_done[3];                 // last work of the job on compute stream k
    cudaEvent_t ev_ctr;                     // counters are back in h_ctr
    bool busy = false;
};

struct PlbContext {
    cudaEvent_t trace_prev_end = nullptr;   // PLB_TRACE: last compute mark of the previous job
    int device;
    cudaStream_t stream;
    cudaStream_t copy_stream;   // H2D of the pipelined host path
    cudaStream_t stream2;       // further compute streams: chunk k of the pipelined host path runs on
    cudaStream_t stream3;       // stream k mod 3, so one chunk's latency-bound tails overlap the next chunks
    cudaStream_t aux_stream;    // D2H of the host path: a chunk's outputs and, last, the job's counters (kept off the
                                // compute streams, where the copies would sit between one job's kernels and the next's)
    cudaEvent_t ev_s2, ev_s3;
    JobSlot slot[kMaxJobs];
    int cur = 0;                // slot whose arena / counters the code below is working with
    int last_done = 0;          // slot of the most recently finished run (plb_last_stats)
    int64_t jobs_submitted = 0;
    bool own_stream;
    int64_t launches;
    int n_sm;
    int smem_optin;
    std::vector<Block> cache;   // free device blocks, reused by size
    cudaEvent_t ev;
    bool timing;
    int n_timed;                                        // runs recorded since plb_set_timing(1)
    cudaEvent_t kev[kTimingRing][kMaxChunks][PLB_N_KERNELS + 1];   // ring of per-run, per-chunk event sets
    int kev_chunks[kTimingRing];                        // chunks recorded in each ring entry
    Counters* d_ctr() { return slot[cur].d_ctr; }
    Counters* h_ctr() { return slot[cur].h_ctr; }
};

struct TileLists {
    std::vector<Tile> a, d;
};

// Launch plan of one chunk of windows [w0, w1): tile lists (host + device) and kernel shapes.
struct ChunkPlan {
    int w0 = 0, w1 = 0;
    int h0 = 0, h1 = 0;   // haplotype range of the chunk
    int r0 = 0, r1 = 0;   // read-index hull of the chunk's slots (k_read_check)
    AnchorPlan ap{};
    DpPlan dp{};
    size_t a_smem = 0, d_smem = 0;
    int a_occ = 1, d_occ = 1;
    Block tiles_blk{nullptr, 0};
};

struct PlbDeviceBatch {
    DevBatch d{};
    Block blk{nullptr, 0};
    Queue q{}, q2{}, q3{};             // general-path queues (one per compute stream)
    Block mode_blk{nullptr, 0};        // larger queues for the run-time modes (every alignment is queued),
    Queue mq{}, mq2{}, mq3{};          // allocated on the first run with calc_flank_score / use_mapq_cap
    double* ll_scratch = nullptr;
    double* em_scratch = nullptr;      // [W][nInd][Gmax_plan]
    int32_t max_haps = 0;              // largest H in the batch
    int32_t max_hap_len = 0;
    int64_t hap_bytes = 0, read_bytes = 0;
    size_t em_scratch_elems = 0;
    bool have_var = false;
    bool shares_reads = false;         // window / slot / read arrays belong to another batch (prepare_batch `share`)
    int64_t n_wi = 0;
    std::vector<ChunkPlan> chunks;     // plb_batch_upload plans one chunk covering every window
    int64_t* h_ll_off = nullptr;       // pinned
    int64_t hap_done[2] = {0, 0}, read_done[2] = {0, 0};  // byte intervals already on the device
    int64_t readmeta_done[2] = {0, 0};                    // read-index interval whose per-read arrays are on the device
    // 2-bit packed input (PLB_SEQ_2BIT): the packed bytes land in pk_*, k_unpack2 restores the ASCII arrays of `d`
    bool packed = false;
    uint8_t* pk_hap = nullptr;
    uint8_t* pk_read = nullptr;
    int64_t n_read_exc = 0, n_hap_exc = 0;
    int64_t* d_read_exc_pos = nullptr;
    uint8_t* d_read_exc_chr = nullptr;
    int64_t* d_hap_exc_pos = nullptr;
    uint8_t* d_hap_exc_chr = nullptr;
    bool exc_uploaded = false;
    int qual_bits = 0;                 // packed qualities: codes land in pk_qual, k_unpack_qual restores d.read_qual
    uint8_t* pk_qual = nullptr;
    QualTable qtab{};
    struct Fresh { int which; int64_t lo, hi; };          // base interval just uploaded: 0 = haplotypes, 1 = reads
    std::vector<Fresh> fresh[kMaxChunks + 1];             // per chunk of the host path ([0] for whole-batch uploads)
};

extern "C" const char* plb_last_error(void) { return g_err; }
extern "C" int plb_abi_version(void) { return PLB_ABI_VERSION; }

extern "C" int plb_context_create(int device, void* stream, PlbContext** out) {
    if (!out) return set_err(PLB_ERR_ARG, "out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_err(PLB_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback",
                       e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return set_err(PLB_ERR_ARG, "device %d out of range (have %d)", device, n);
    CU(cudaSetDevice(device));
    {   // log(1 - exp(mLTOT * mapq)) with the host's libm (see c_map_right); mapq 0 gives log(0) = -inf, as in the reference
        double right[256];
        for (int m = 0; m < 256; ++m) right[m] = log(1.0 - exp(kMLTOT * (double)m));
        CU(cudaMemcpyToSymbol(c_map_right, right, sizeof right));
    }
    PlbContext* c = new PlbContext();
    c->device = device;
    c->launches = 0;
    if (stream) {
        c->stream = (cudaStream_t)stream;
        c->own_stream = false;
    } else {
        CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->n_sm = prop.multiProcessorCount;
    c->smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (prop.major < 10)
        return set_err(PLB_ERR_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", device,
                       prop.major, prop.minor);
    for (int j = 0; j < kMaxJobs; ++j) {
        JobSlot& js = c->slot[j];
        CU(cudaMalloc(&js.d_ctr, sizeof(Counters)));
        CU(cudaMallocHost(&js.h_ctr, sizeof(Counters)));
        memset(js.h_ctr, 0, sizeof(Counters));
        for (int i = 0; i <= kMaxChunks; ++i) CU(cudaEventCreateWithFlags(&js.ev_chunk[i], cudaEventDisableTiming));
        for (int i = 0; i < kMaxChunks; ++i) CU(cudaEventCreateWithFlags(&js.ev_unpack[i], cudaEventDisableTiming));
        for (int i = 0; i < kMaxChunks; ++i) CU(cudaEventCreateWithFlags(&js.ev_tiles[i], cudaEventDisableTiming));
        for (int i = 0; i < kMaxChunks; ++i) CU(cudaEventCreateWithFlags(&js.ev_anchor[i], cudaEventDisableTiming));
        for (int i = 0; i < kMaxChunks; ++i) CU(cudaEventCreateWithFlags(&js.ev_kdone[i], cudaEventDisableTiming));
        for (int i = 0; i < 3; ++i) CU(cudaEventCreateWithFlags(&js.ev_done[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&js.ev_ctr, cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&c->ev, cudaEventDisableTiming));
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_s2, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_s3, cudaEventDisableTiming));
    c->timing = false;
    c->n_timed = 0;
    for (int r = 0; r < kTimingRing; ++r)
        for (int k = 0; k < kMaxChunks; ++k)
            for (int i = 0; i <= PLB_N_KERNELS; ++i) CU(cudaEventCreate(&c->kev[r][k][i]));
    *out = c;
    return PLB_OK;
}

extern "C" void plb_context_destroy(PlbContext* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamSynchronize(c->stream2);
    cudaStreamSynchronize(c->stream3);
    cudaStreamSynchronize(c->aux_stream);
    for (auto& b : c->cache) cudaFree(b.p);
    for (int j = 0; j < kMaxJobs; ++j) {
        JobSlot& js = c->slot[j];
        cudaFree(js.d_ctr);
        cudaFreeHost(js.h_ctr);
        for (auto& pb : js.pin.blocks) cudaFreeHost(pb.first);
        for (int i = 0; i <= kMaxChunks; ++i) cudaEventDestroy(js.ev_chunk[i]);
        for (int i = 0; i < kMaxChunks; ++i) cudaEventDestroy(js.ev_unpack[i]);
        for (int i = 0; i < kMaxChunks; ++i) cudaEventDestroy(js.ev_tiles[i]);
        for (int i = 0; i < kMaxChunks; ++i) cudaEventDestroy(js.ev_anchor[i]);
        for (int i = 0; i < kMaxChunks; ++i) cudaEventDestroy(js.ev_kdone[i]);
        for (int i = 0; i < 3; ++i) cudaEventDestroy(js.ev_done[i]);
        cudaEventDestroy(js.ev_ctr);
    }
    cudaEventDestroy(c->ev);
    cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->stream2);
    cudaStreamDestroy(c->stream3);
    cudaStreamDestroy(c->aux_stream);
    cudaEventDestroy(c->ev_s2);
    cudaEventDestroy(c->ev_s3);
    for (int r = 0; r < kTimingRing; ++r)
        for (int k = 0; k < kMaxChunks; ++k)
            for (int i = 0; i <= PLB_N_KERNELS; ++i) cudaEventDestroy(c->kev[r][k][i]);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int64_t plb_launch_count(const PlbContext* c) { return c ? c->launches : 0; }

static int block_get(PlbContext* c, size_t bytes, Block* out) {
    int best = -1;
    for (size_t i = 0; i < c->cache.size(); ++i)
        if (c->cache[i].cap >= bytes && (best < 0 || c->cache[i].cap < c->cache[best].cap)) best = (int)i;
    if (best >= 0) {
        *out = c->cache[best];
        c->cache.erase(c->cache.begin() + best);
        return PLB_OK;
    }
    // drop smaller cached blocks before growing
    for (auto& b : c->cache) cudaFree(b.p);
    c->cache.clear();
    size_t cap = bytes + bytes / 8 + (1 << 20);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, cap);
    if (e != cudaSuccess) return set_err(PLB_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(e));
    out->p = p;
    out->cap = cap;
    return PLB_OK;
}

static void block_put(PlbContext* c, Block b) {
    if (b.p) c->cache.push_back(b);
}

// Pinned host staging memory of the current job slot (valid until the slot's next pin_reset: a slot is reused only after
// its job has been waited for, so nothing in flight can still read it).
static void pin_reset(PlbContext* c) { c->slot[c->cur].pin.off = 0; }

static void* pin_alloc(PlbContext* c, size_t bytes) {
    PinArena& A = c->slot[c->cur].pin;
    bytes = (bytes + 255) & ~(size_t)255;
    if (!A.blocks.empty()) {
        auto& last = A.blocks.back();
        if (A.off + bytes <= last.second) {
            void* p = last.first + A.off;
            A.off += bytes;
            return p;
        }
    }
    // grow: a new block twice the size; older blocks stay alive (copies from them may be in flight)
    size_t cap = std::max<size_t>(bytes * 2, A.blocks.empty() ? (size_t)(4 << 20) : A.blocks.back().second * 2);
    void* p = nullptr;
    if (cudaHostAlloc(&p, cap, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    A.blocks.push_back({(uint8_t*)p, cap});
    A.off = bytes;
    return p;
}

// Entry points that use the context's streams synchronously (device-resident runs, uploads, the selection loop, the S1
// batches) must not interleave with host-path jobs that are still in flight.
static int require_idle(PlbContext* c, const char* what) {
    for (int j = 0; j < kMaxJobs; ++j)
        if (c->slot[j].busy)
            return set_err(PLB_ERR_ARG, "%s: a submitted job is still in flight on this context (plb_population_wait it first)", what);
    c->cur = 0;
    return PLB_OK;
}

// ---- host helpers -----------------------------------------------------------------------------

extern "C" int plb_ll_offsets(const PlbWindowBatch* b, int64_t* ll_off, int64_t* total) {
    if (!b || !ll_off) return set_err(PLB_ERR_ARG, "NULL argument");
    int64_t tot = 0;
    for (int w = 0; w < b->n_windows; ++w)
        for (int i = 0; i < b->n_individuals; ++i) {
            int64_t wi = (int64_t)w * b->n_individuals + i;
            ll_off[wi] = tot;
            tot += (int64_t)(b->win_hap_off[w + 1] - b->win_hap_off[w]) * (b->wi_slot_off[wi + 1] - b->wi_slot_off[wi]);
        }
    ll_off[(int64_t)b->n_windows * b->n_individuals] = tot;
    if (total) *total = tot;
    return PLB_OK;
}

static int check_options(const PlbOptions* o) {
    if (!o) return set_err(PLB_ERR_ARG, "options is NULL");
    if (o->use_mapq_cap != 0 && o->use_mapq_cap != 1) return set_err(PLB_ERR_ARG, "use_mapq_cap must be 0 or 1");
    if (o->calc_flank_score != 0 && o->calc_flank_score != 1)
        return set_err(PLB_ERR_ARG, "calc_flank_score must be 0 or 1");
    if (o->gap_extend < 0 || o->gap_extend > 64 || o->nuc_prior < 0 || o->nuc_prior > 64)
        return set_err(PLB_ERR_ARG, "gap_extend / nuc_prior out of range");
    return PLB_OK;
}

// The read-length rule applies to the reads the kernels will actually score: good / bad reads that fail QC or overlap
// the window by fewer than 7 bases are short-circuited to LL = 0 (chaplotype.pyx:343-361) exactly as in k_anchor.
static bool slot_is_scored(const PlbWindowBatch* b, int w, int64_t wi, int64_t s, int r) {
    const int64_t t = s - b->wi_slot_off[wi];
    if (t >= (int64_t)b->wi_n_good[wi] + b->wi_n_bad[wi]) return true;   // broken mates are always scored
    if (b->read_qcfail[r]) return false;
    const int ws = b->win_start[w], we = b->win_end[w], rp = b->read_pos[r], re = b->read_end[r];
    const int lo = ws > rp ? ws : rp, hi = we < re ? we : re;
    return hi > lo && hi - lo >= PLB_KMER;
}

// Checks a HOST batch.  O(windows + haplotypes + slots), spread over the library's host threads; `bytes` adds the scan
// of every base quality (the run paths leave that to the kernels, which flag it while they build their profiles).
static int validate_batch(const PlbWindowBatch* b, const PlbOptions* opt, int32_t max_haps, bool bytes) {
    if (!b) return set_err(PLB_ERR_ARG, "batch is NULL");
    if (opt) {
        int rc = check_options(opt);
        if (rc) return rc;
    }
    if (b->n_windows < 0 || b->n_individuals < 1) return set_err(PLB_ERR_ARG, "bad n_windows / n_individuals");
    if (b->n_windows == 0) return PLB_OK;
    if (!b->win_hap_off || !b->win_start || !b->win_end || !b->hap_start || !b->hap_seq_off || !b->wi_slot_off ||
        !b->wi_n_good || !b->wi_n_bad || !b->read_seq_off)
        return set_err(PLB_ERR_ARG, "NULL array in batch");
    if (b->win_hap_off[0] != 0 || b->win_hap_off[b->n_windows] != b->n_haps)
        return set_err(PLB_ERR_ARG, "win_hap_off inconsistent with n_haps");
    const int64_t nwi = (int64_t)b->n_windows * b->n_individuals;
    if (b->wi_slot_off[0] != 0 || b->wi_slot_off[nwi] != b->n_slots)
        return set_err(PLB_ERR_ARG, "wi_slot_off inconsistent with n_slots");
    if (b->n_slots > 0 && (!b->slot_read || !b->read_seq || !b->read_qual || !b->read_pos || !b->read_end ||
                           !b->read_mapq || !b->read_qcfail))
        return set_err(PLB_ERR_ARG, "NULL read array in batch");
    if (b->n_haps > 0 && !b->hap_seq) return set_err(PLB_ERR_ARG, "hap_seq is NULL");
    if (b->max_variants > 64) return set_err(PLB_ERR_SHAPE, "max_variants %d > 64", b->max_variants);
    if (b->seq_format != PLB_SEQ_ASCII && b->seq_format != PLB_SEQ_2BIT)
        return set_err(PLB_ERR_ARG, "seq_format %d unknown", b->seq_format);
    if (b->qual_bits != 0 && b->qual_bits != 4 && b->qual_bits != 6) return set_err(PLB_ERR_ARG, "qual_bits %d: must be 0, 4 or 6", b->qual_bits);
    for (int i = 0; i < (b->qual_bits ? 1 << b->qual_bits : 0); ++i)
        if (b->qual_table[i] > 93) return set_err(PLB_ERR_ARG, "qual_table[%d] = %d > 93", i, b->qual_table[i]);
    if (b->seq_format == PLB_SEQ_2BIT) {
        if (b->n_read_exc < 0 || b->n_hap_exc < 0 || (b->n_read_exc > 0 && (!b->read_exc_pos || !b->read_exc_chr)) ||
            (b->n_hap_exc > 0 && (!b->hap_exc_pos || !b->hap_exc_chr)))
            return set_err(PLB_ERR_ARG, "packed batch: bad exception lists");
        const int64_t rb = b->n_reads ? b->read_seq_off[b->n_reads] : 0, hbts = b->n_haps ? b->hap_seq_off[b->n_haps] : 0;
        for (int64_t i = 0; i < b->n_read_exc; ++i)
            if (b->read_exc_pos[i] < 0 || b->read_exc_pos[i] >= rb)
                return set_err(PLB_ERR_ARG, "packed batch: read exception %lld out of range", (long long)i);
        for (int64_t i = 0; i < b->n_hap_exc; ++i)
            if (b->hap_exc_pos[i] < 0 || b->hap_exc_pos[i] >= hbts)
                return set_err(PLB_ERR_ARG, "packed batch: haplotype exception %lld out of range", (long long)i);
    }
    // first failing window wins, as in a serial scan (messages name the lowest window)
    int bad_w = b->n_windows;
    int bad_code = PLB_OK;
    char bad_msg[400] = "";
    const bool hla = opt && opt->use_mapq_cap;
#pragma omp parallel for schedule(static, 256) num_threads(host_threads()) if (b->n_windows > 2048)
    for (int w = 0; w < b->n_windows; ++w) {
        if (w > bad_w) continue;
        char msg[400] = "";
        int code = PLB_OK;
        auto fail = [&](int cde, const char* fmt, ...) {
            if (code != PLB_OK) return;
            code = cde;
            va_list ap;
            va_start(ap, fmt);
            vsnprintf(msg, sizeof msg, fmt, ap);
            va_end(ap);
        };
        const int H = b->win_hap_off[w + 1] - b->win_hap_off[w];
        if (H < 1) fail(PLB_ERR_SHAPE, "window %d has no haplotypes", w);
        if (max_haps > 0 && H > max_haps)
            fail(PLB_ERR_SHAPE, "window %d has %d haplotypes > max_haps %d (cpopulation.pyx:221)", w, H, max_haps);
        if (opt && opt->calc_flank_score && b->win_start[w] - b->hap_start[w] <= 0)
            fail(PLB_ERR_ARG, "window %d: calc_flank_score needs a positive flank (win_start - hap_start)", w);
        int min_len = 1 << 30;
        for (int h = b->win_hap_off[w]; h < b->win_hap_off[w + 1] && code == PLB_OK; ++h) {
            const int64_t len = b->hap_seq_off[h + 1] - b->hap_seq_off[h];
            if (len < 0) fail(PLB_ERR_ARG, "hap_seq_off not monotone at %d", h);
            if (len > PLB_MAX_HAP_LEN)
                fail(PLB_ERR_SHAPE, "haplotype %d is %lld bp > %d (chaplotype.pyx:180)", h, (long long)len, PLB_MAX_HAP_LEN);
            min_len = std::min<int>(min_len, (int)len);
        }
        for (int i = 0; i < b->n_individuals && code == PLB_OK; ++i) {
            const int64_t wi = (int64_t)w * b->n_individuals + i;
            const int64_t T = b->wi_slot_off[wi + 1] - b->wi_slot_off[wi];
            if (T < 0 || b->wi_n_good[wi] < 0 || b->wi_n_bad[wi] < 0 || b->wi_n_good[wi] + b->wi_n_bad[wi] > T) {
                fail(PLB_ERR_ARG, "read counts inconsistent for window %d individual %d", w, i);
                break;
            }
            for (int64_t s = b->wi_slot_off[wi]; s < b->wi_slot_off[wi + 1] && code == PLB_OK; ++s) {
                const int r = b->slot_read[s];
                if (r < 0 || r >= b->n_reads) {
                    fail(PLB_ERR_ARG, "slot %lld: read index out of range", (long long)s);
                    break;
                }
                const int64_t L = b->read_seq_off[r + 1] - b->read_seq_off[r];
                if (L < 0 || L > 32767) fail(PLB_ERR_SHAPE, "read %d length %lld out of range", r, (long long)L);
                if (code != PLB_OK || L < PLB_KMER || !slot_is_scored(b, w, wi, s, r)) continue;
                // the reference would read past the haplotype (calign.pyx:256-259); refuse instead
                if (!hla && L + 15 > min_len)
                    fail(PLB_ERR_SHAPE, "window %d: read %d (%lld bp) + 15 exceeds haplotype length %d", w, r, (long long)L,
                         min_len);
                if (hla) {  // same rule on the read as clipped to each haplotype (chaplotype.pyx:647-655)
                    for (int h = b->win_hap_off[w]; h < b->win_hap_off[w + 1]; ++h) {
                        const int hl = (int)(b->hap_seq_off[h + 1] - b->hap_seq_off[h]);
                        const int o1 = std::max(0, b->hap_start[w] - b->read_pos[r]);
                        const int o2 = std::max(0, b->read_pos[r] + (int)L - b->win_start[w] - hl);
                        const int Lc = (int)L - o1 - o2;
                        if (Lc >= PLB_KMER && Lc + 15 > hl)
                            fail(PLB_ERR_SHAPE, "window %d: clipped read %d (%d bp) + 15 exceeds haplotype length %d", w, r, Lc, hl);
                    }
                }
            }
        }
        if (code != PLB_OK) {
#pragma omp critical(plb_validate_err)
            if (w < bad_w) {
                bad_w = w;
                bad_code = code;
                memcpy(bad_msg, msg, sizeof bad_msg);
            }
        }
    }
    if (bad_code != PLB_OK) return set_err(bad_code, "%s", bad_msg);
    if (bytes && b->qual_bits == 0) {   // packed qualities: the table was checked above
        const int64_t nb = b->n_reads ? b->read_seq_off[b->n_reads] : 0;
        int64_t bad = -1;
#pragma omp parallel for schedule(static) num_threads(host_threads()) reduction(max : bad) if (nb > (1 << 20))
        for (int64_t i = 0; i < nb; ++i)
            if (b->read_qual[i] > 93) bad = std::max(bad, i);
        if (bad >= 0) return set_err(PLB_ERR_ARG, "base quality %d > 93 at byte %lld", b->read_qual[bad], (long long)bad);
    }
    return PLB_OK;
}

extern "C" int plb_validate(const PlbWindowBatch* b, const PlbOptions* opt, int32_t max_haps) {
    return validate_batch(b, opt, max_haps, true);
}

// ---- upload + planning ----------------------------------------------------------------------

namespace {

struct Layout {  // bump allocator over one device block
    size_t off = 0;
    size_t take(size_t bytes) {
        size_t o = off;
        off = (off + bytes + 255) & ~(size_t)255;
        return o;
    }
};

template <typename T>
T* at(const Block& b, size_t off) {
    return (T*)((uint8_t*)b.p + off);
}

// position chains of one haplotype group.  The chain heads and multiplicity rows grow with the group as well (about
// 6 bytes per haplotype base in all), so 12 KB of chains keeps a tile near 45 KB of shared memory = 4 CTAs per SM; with
// 24 KB the 48-haplotype groups of the selection rounds needed 98 KB (2 CTAs per SM, 25 % of the warps active).
// Measured on the selection workload (k_anchor per round launch): 24 KB 0.82 ms, 12 KB 0.58, 8 KB 0.60, 6 KB 0.63, 4 KB 0.69.
// PLB_ANCHOR_BUDGET overrides it for such sweeps.
static const size_t kAnchorNextBudget = getenv("PLB_ANCHOR_BUDGET") ? (size_t)atoi(getenv("PLB_ANCHOR_BUDGET")) : 12 * 1024;
constexpr size_t kAnchorHashBudget = 24 * 1024;  // read 7-mer ids of one tile
constexpr size_t kAnchorCntBudget = 48 * 1024;   // per-warp vote arrays of the exact (tie) path
constexpr int kAnchorMaxSlots = 256;
constexpr int kAnchorMaxPairs = 4096;
constexpr size_t kDpRecBudget = 40 * 1024;
constexpr size_t kDpProfBudget = 56 * 1024;
constexpr int kDpMaxSlots = 128;
constexpr int kDpMaxPairs = 4096;
constexpr int kDpThreads = 256;

int prof_row_words(int L) {
    int n = dp_steps(L) + 4;
    n = (n + 3) & ~3;
    if (!((n >> 2) & 1)) n += 4;
    return n;
}

}  // namespace

static int plan_tiles(const PlbWindowBatch* hb, int w_begin, int w_end, TileLists& tl, AnchorPlan& ap, DpPlan& dp,
                      int& max_read, int& max_hap, int& max_H) {
    memset(&ap, 0, sizeof ap);
    memset(&dp, 0, sizeof dp);
    max_read = 0;
    max_hap = 0;
    max_H = 0;
    const int nInd = hb->n_individuals;
    std::vector<int> slot_len;
    std::vector<std::pair<int, int>> groups, dgroups;
    for (int w = w_begin; w < w_end; ++w) {
        groups.clear();
        dgroups.clear();
        const int h0 = hb->win_hap_off[w], h1 = hb->win_hap_off[w + 1];
        max_H = std::max(max_H, h1 - h0);
        const int64_t s0 = hb->wi_slot_off[(int64_t)w * nInd], s1 = hb->wi_slot_off[(int64_t)(w + 1) * nInd];
        slot_len.resize((size_t)(s1 - s0));
        for (int64_t s = s0; s < s1; ++s) {
            const int r = hb->slot_read[s];
            slot_len[(size_t)(s - s0)] = (int)(hb->read_seq_off[r + 1] - hb->read_seq_off[r]);
            max_read = std::max(max_read, slot_len[(size_t)(s - s0)]);
        }
        // ---- anchor tiles
        {
            // greedy grouping under the budget, then the same number of groups with the haplotypes spread evenly
            // (50 trial haplotypes: 13 + 13 + 13 + 11 rather than 16 + 16 + 16 + 2)
            auto build = [&](int cap) {
                groups.clear();
                int g0 = h0;
                size_t next_bytes = 0;
                for (int h = h0; h < h1; ++h) {
                    const int len = (int)(hb->hap_seq_off[h + 1] - hb->hap_seq_off[h]);
                    max_hap = std::max(max_hap, len);
                    const size_t need = 2 * (size_t)((len + 2) & ~1);
                    if (h > g0 && (next_bytes + need > kAnchorNextBudget || h - g0 >= cap)) {
                        groups.push_back({g0, h});
                        g0 = h;
                        next_bytes = 0;
                    }
                    next_bytes += need;
                }
                groups.push_back({g0, h1});
            };
            build(64);
            const int ng = (int)groups.size();
            if (ng > 1) {
                build((h1 - h0 + ng - 1) / ng);
                if ((int)groups.size() > ng) build(64);   // uneven lengths: the even split needed more groups
            }
            for (auto& g : groups) {
                int nh_ = 0, sum_nk = 0, hpk = 0;
                for (int h = g.first; h < g.second; ++h) {
                    const int len = (int)(hb->hap_seq_off[h + 1] - hb->hap_seq_off[h]);
                    nh_ += (len + 2) & ~1;
                    sum_nk += std::max(0, len - kKmer);
                    hpk += ((len + 15) >> 4) + 2 * kPackPadWords;
                }
                ap.hpk_words = std::max(ap.hpk_words, hpk);
                ap.next_halfs = std::max(ap.next_halfs, nh_);
                ap.max_group = std::max(ap.max_group, g.second - g.first);
                ap.heads_halfs = std::max(ap.heads_halfs, std::min(kHashSize, sum_nk) + 1);
            }
        }
        {
            int maxg = 1;
            for (auto& g : groups) maxg = std::max(maxg, g.second - g.first);
            int64_t c0 = s0;
            size_t halfs = 0, pkw = 0;
            auto flush = [&](int64_t c1) {
                if (c1 <= c0) return;
                for (auto& g : groups) {
                    tl.a.push_back(Tile{w, g.first, g.second, c0, c1});
                    ap.max_pairs = std::max<int>(ap.max_pairs, (int)(c1 - c0) * (g.second - g.first));
                }
                ap.max_slots = std::max<int>(ap.max_slots, (int)(c1 - c0));
                ap.rpk_words = std::max<int>(ap.rpk_words, (int)pkw);
            };
            for (int64_t s = s0; s < s1; ++s) {
                const int len = slot_len[(size_t)(s - s0)];
                const int nk = (std::max(0, len - kKmer) + 7) & ~7;
                if (s > c0 && (2 * (halfs + nk) > kAnchorHashBudget || s - c0 >= kAnchorMaxSlots ||
                               (s - c0 + 1) * maxg > kAnchorMaxPairs)) {
                    flush(s);
                    c0 = s;
                    halfs = 0;
                    pkw = 0;
                }
                halfs += nk;
                pkw += ((len + 15) >> 4) + kPackPadWords;
            }
            flush(s1);
        }
        // ---- dp tiles
        {
            int g0 = h0;
            size_t recs = 0;
            for (int h = h0; h < h1; ++h) {
                const size_t need = (size_t)(hb->hap_seq_off[h + 1] - hb->hap_seq_off[h]) + kRecPad;
                if (h > g0 && ((recs + need) * sizeof(HapRec) > kDpRecBudget || h - g0 >= 64)) {
                    dgroups.push_back({g0, h});
                    g0 = h;
                    recs = 0;
                }
                recs += need;
            }
            dgroups.push_back({g0, h1});
            for (auto& g : dgroups) {
                size_t r = 0;
                for (int h = g.first; h < g.second; ++h) r += (size_t)(hb->hap_seq_off[h + 1] - hb->hap_seq_off[h]) + kRecPad;
                dp.rec_count = std::max<int>(dp.rec_count, (int)r);
                dp.max_group = std::max(dp.max_group, g.second - g.first);
            }
        }
        {
            int maxg = 1;
            for (auto& g : dgroups) maxg = std::max(maxg, g.second - g.first);
            int64_t c0 = s0;
            size_t words = 0;
            auto flush = [&](int64_t c1) {
                if (c1 <= c0) return;
                for (auto& g : dgroups) {
                    tl.d.push_back(Tile{w, g.first, g.second, c0, c1});
                    dp.max_pairs = std::max<int>(dp.max_pairs, (int)(c1 - c0) * (g.second - g.first));
                }
                dp.max_slots = std::max<int>(dp.max_slots, (int)(c1 - c0));
                dp.prof_words = std::max<int>(dp.prof_words, (int)words);
            };
            for (int64_t s = s0; s < s1; ++s) {
                const int L = slot_len[(size_t)(s - s0)];
                const int wds = (L >= kMinFastLen && L <= kMaxFastLen) ? prof_row_words(L) : 0;
                if (s > c0 && (4 * (words + wds) > kDpProfBudget || s - c0 >= kDpMaxSlots ||
                               (s - c0 + 1) * maxg > kDpMaxPairs)) {
                    flush(s);
                    c0 = s;
                    words = 0;
                }
                words += wds;
            }
            flush(s1);
        }
    }
    ap.n_tiles = (int)tl.a.size();
    dp.n_tiles = (int)tl.d.size();
    ap.max_slots = std::max(ap.max_slots, 1);
    ap.max_group = std::max(ap.max_group, 1);
    dp.max_slots = std::max(dp.max_slots, 1);
    dp.max_group = std::max(dp.max_group, 1);
    dp.max_pairs = std::max(dp.max_pairs, 1);
    return PLB_OK;
}

// Bytes of the big sequence arrays (hap_seq / read_seq / read_qual) that windows [w0, w1) need.
struct ByteRanges {
    int64_t hap0, hap1;    // byte range of hap_seq
    int64_t read0, read1;  // byte range of read_seq / read_qual (hull over the reads the slots refer to)
    int rmin, rmax;        // that hull as read indices (rmax < rmin: no reads)
};

static ByteRanges byte_ranges(const PlbWindowBatch* hb, int w0, int w1) {
    ByteRanges r{0, 0, 0, 0, 0, -1};
    if (w1 <= w0) return r;
    const int nInd = hb->n_individuals;
    r.hap0 = hb->hap_seq_off[hb->win_hap_off[w0]];
    r.hap1 = hb->hap_seq_off[hb->win_hap_off[w1]];
    const int64_t s0 = hb->wi_slot_off[(int64_t)w0 * nInd], s1 = hb->wi_slot_off[(int64_t)w1 * nInd];
    int rmin = INT32_MAX, rmax = -1;
    const int32_t* sr = hb->slot_read;
#pragma omp parallel for reduction(min : rmin) reduction(max : rmax) schedule(static) num_threads(host_threads()) \
    if (s1 - s0 > 32768)
    for (int64_t s = s0; s < s1; ++s) {
        const int rd = sr[s];
        rmin = std::min(rmin, rd);
        rmax = std::max(rmax, rd);
    }
    if (rmax >= 0) {
        r.read0 = hb->read_seq_off[rmin];
        r.read1 = hb->read_seq_off[rmax + 1];
        r.rmin = rmin;
        r.rmax = rmax;
    }
    return r;
}

// Plans tiles, allocates one device block and lays every array out in it.  No copies.
// `share` (the selection rounds, plb_select.cuh): the batch covers a prefix of the windows of `share` with the very
// same slots and reads but other haplotypes; the window-, slot- and read-side arrays are then those of `share`
// (already on the device) and only the haplotype-side arrays and the scratch are laid out.
static int prepare_batch(PlbContext* c, const PlbWindowBatch* hb, PlbDeviceBatch** out,
                         const PlbDeviceBatch* share = nullptr, bool reset_arena = true) {
    if (!c || !hb || !out) return set_err(PLB_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    const int W = hb->n_windows, nInd = hb->n_individuals;
    if (W < 0 || nInd < 1) return set_err(PLB_ERR_ARG, "bad n_windows / n_individuals");
    const int64_t nwi = (int64_t)W * nInd;
    const int64_t n_slots = hb->n_slots;
    const int n_haps = hb->n_haps, n_reads = hb->n_reads;
    const int64_t hap_bytes = n_haps ? hb->hap_seq_off[n_haps] : 0;
    const int64_t read_bytes = n_reads ? hb->read_seq_off[n_reads] : 0;

    PlbDeviceBatch* db = new PlbDeviceBatch();
    db->n_wi = nwi;
    if (!share && reset_arena) pin_reset(c);   // the rounds of a selection keep the arena of their call
    db->h_ll_off = (int64_t*)pin_alloc(c, ((size_t)nwi + 1) * 8);
    if (!db->h_ll_off) {
        delete db;
        return set_err(PLB_ERR_NOMEM, "pinned host allocation failed");
    }
    int64_t n_pairs = 0;
    plb_ll_offsets(hb, db->h_ll_off, &n_pairs);

    int max_H = 0;
    for (int w = 0; w < W; ++w) max_H = std::max(max_H, hb->win_hap_off[w + 1] - hb->win_hap_off[w]);
    db->max_haps = max_H;
    for (int h = 0; h < n_haps; ++h) db->max_hap_len = std::max<int32_t>(db->max_hap_len, (int32_t)(hb->hap_seq_off[h + 1] - hb->hap_seq_off[h]));
    db->hap_bytes = hap_bytes;
    db->read_bytes = read_bytes;

    // device layout
    Layout L;
    const size_t own = share ? 0 : 1;   // 0: the array is the one of `share`
    const size_t o_win_hap_off = L.take((size_t)(W + 1) * 4), o_win_start = L.take(own * W * 4),
                 o_win_end = L.take(own * W * 4), o_hap_start = L.take(own * W * 4),
                 o_hap_seq_off = L.take((size_t)(n_haps + 1) * 8), o_hap_seq = L.take((size_t)hap_bytes + 64),
                 o_wi_slot_off = L.take(own * (nwi + 1) * 8), o_wi_n_good = L.take(own * nwi * 4),
                 o_wi_n_bad = L.take(own * nwi * 4), o_slot_read = L.take(own * n_slots * 4),
                 o_read_seq_off = L.take(own * (n_reads + 1) * 8), o_read_seq = L.take(own * (read_bytes + 64)),
                 o_read_qual = L.take(own * (read_bytes + 64)), o_read_pos = L.take(own * n_reads * 4),
                 o_read_end = L.take(own * n_reads * 4), o_read_mapq = L.take(own * n_reads),
                 o_read_qcfail = L.take(own * n_reads);
    const bool have_var = hb->max_variants > 0 && hb->win_n_var && hb->hap_var_mask && hb->var_prior;
    db->have_var = have_var;
    const size_t o_win_n_var = L.take(have_var ? (size_t)W * 4 : 0), o_hap_var_mask = L.take(have_var ? (size_t)n_haps * 8 : 0),
                 o_var_prior = L.take(have_var ? (size_t)W * hb->max_variants * 8 : 0);
    const bool packed = hb->seq_format == PLB_SEQ_2BIT;
    if (packed && share) {
        delete db;
        return set_err(PLB_ERR_UNSUPPORTED, "packed batches are not taken by the selection rounds");
    }
    db->packed = packed;
    db->n_read_exc = packed ? hb->n_read_exc : 0;
    db->n_hap_exc = packed ? hb->n_hap_exc : 0;
    const int qbits = hb->qual_bits;
    if (qbits && share) {
        delete db;
        return set_err(PLB_ERR_UNSUPPORTED, "packed batches are not taken by the selection rounds");
    }
    db->qual_bits = qbits;
    if (qbits) memcpy(db->qtab.w, hb->qual_table, 64);
    const size_t o_pk_qual = L.take(qbits ? (size_t)(read_bytes * qbits + 7) / 8 + 64 : 0);
    const size_t o_pk_hap = L.take(packed ? (size_t)(hap_bytes + 3) / 4 + 64 : 0),
                 o_pk_read = L.take(packed ? (size_t)(read_bytes + 3) / 4 + 64 : 0),
                 o_rexc_pos = L.take((size_t)db->n_read_exc * 8), o_rexc_chr = L.take((size_t)db->n_read_exc),
                 o_hexc_pos = L.take((size_t)db->n_hap_exc * 8), o_hexc_chr = L.take((size_t)db->n_hap_exc);
    const size_t o_slot_wi = L.take(own * n_slots * 4), o_hap_win = L.take((size_t)n_haps * 4),
                 o_ll_off = L.take((size_t)(nwi + 1) * 8);
    const size_t o_rflags = L.take(own * ((size_t)n_reads + 64));
    const size_t o_gap = L.take((size_t)hap_bytes + n_haps + 64), o_wgen = L.take((size_t)W * 4 + 64),
                 o_c0 = L.take((size_t)n_pairs * 4), o_c1 = L.take((size_t)n_pairs * 4),
                 o_score = L.take((size_t)n_pairs * 4);
    const int qcap = (int)std::min<int64_t>(std::max<int64_t>(4096, n_pairs / 2), 1 << 26);
    const size_t o_q = L.take((size_t)qcap * sizeof(QueueEntry)), o_qcount = L.take(64);
    const size_t o_q2 = L.take((size_t)qcap * sizeof(QueueEntry)), o_qcount2 = L.take(64);
    const size_t o_q3 = L.take((size_t)qcap * sizeof(QueueEntry)), o_qcount3 = L.take(64);
    const size_t o_ll = L.take((size_t)n_pairs * 8);
    const int Gp = max_H * (max_H + 1) / 2;
    db->em_scratch_elems = share ? 0 : (size_t)W * nInd * Gp;   // the selection rounds never run the window model
    const size_t o_em = L.take(db->em_scratch_elems * 8);

    int rc = block_get(c, L.off + 256, &db->blk);
    if (rc) {
        delete db;
        return rc;
    }
    const Block& B = db->blk;
    DevBatch& d = db->d;
    d.n_windows = W;
    d.n_individuals = nInd;
    d.n_haps = n_haps;
    d.n_reads = n_reads;
    d.n_slots = n_slots;
    d.n_pairs = n_pairs;
    d.win_hap_off = at<int32_t>(B, o_win_hap_off);
    d.win_start = at<int32_t>(B, o_win_start);
    d.win_end = at<int32_t>(B, o_win_end);
    d.hap_start = at<int32_t>(B, o_hap_start);
    d.hap_seq_off = at<int64_t>(B, o_hap_seq_off);
    d.hap_seq = at<uint8_t>(B, o_hap_seq);
    d.wi_slot_off = at<int64_t>(B, o_wi_slot_off);
    d.wi_n_good = at<int32_t>(B, o_wi_n_good);
    d.wi_n_bad = at<int32_t>(B, o_wi_n_bad);
    d.slot_read = at<int32_t>(B, o_slot_read);
    d.read_seq_off = at<int64_t>(B, o_read_seq_off);
    d.read_seq = at<uint8_t>(B, o_read_seq);
    d.read_qual = at<uint8_t>(B, o_read_qual);
    d.read_pos = at<int32_t>(B, o_read_pos);
    d.read_end = at<int32_t>(B, o_read_end);
    d.read_mapq = at<uint8_t>(B, o_read_mapq);
    d.read_qcfail = at<uint8_t>(B, o_read_qcfail);
    d.max_variants = have_var ? hb->max_variants : 0;
    d.win_n_var = have_var ? at<int32_t>(B, o_win_n_var) : nullptr;
    d.hap_var_mask = have_var ? at<uint64_t>(B, o_hap_var_mask) : nullptr;
    d.var_prior = have_var ? at<double>(B, o_var_prior) : nullptr;
    d.slot_wi = at<int32_t>(B, o_slot_wi);
    d.hap_win = at<int32_t>(B, o_hap_win);
    d.ll_off = at<int64_t>(B, o_ll_off);
    d.gap_open = at<uint8_t>(B, o_gap);
    d.win_flags = at<uint32_t>(B, o_wgen);
    d.cand0 = at<int32_t>(B, o_c0);
    d.cand1 = at<int32_t>(B, o_c1);
    d.score = at<int32_t>(B, o_score);
    d.read_flags = at<uint8_t>(B, o_rflags);
    db->q.e = at<QueueEntry>(B, o_q);
    db->q.count = at<int32_t>(B, o_qcount);
    db->q.cap = qcap;
    db->q2.e = at<QueueEntry>(B, o_q2);
    db->q2.count = at<int32_t>(B, o_qcount2);
    db->q2.cap = qcap;
    db->q3.e = at<QueueEntry>(B, o_q3);
    db->q3.count = at<int32_t>(B, o_qcount3);
    db->q3.cap = qcap;
    db->ll_scratch = at<double>(B, o_ll);
    db->em_scratch = at<double>(B, o_em);
    if (qbits) db->pk_qual = at<uint8_t>(B, o_pk_qual);
    if (packed) {
        db->pk_hap = at<uint8_t>(B, o_pk_hap);
        db->pk_read = at<uint8_t>(B, o_pk_read);
        db->d_read_exc_pos = at<int64_t>(B, o_rexc_pos);
        db->d_read_exc_chr = at<uint8_t>(B, o_rexc_chr);
        db->d_hap_exc_pos = at<int64_t>(B, o_hexc_pos);
        db->d_hap_exc_chr = at<uint8_t>(B, o_hexc_chr);
    }
    if (share) {
        const DevBatch& s = share->d;
        d.win_start = s.win_start;
        d.win_end = s.win_end;
        d.hap_start = s.hap_start;
        d.wi_slot_off = s.wi_slot_off;
        d.wi_n_good = s.wi_n_good;
        d.wi_n_bad = s.wi_n_bad;
        d.slot_read = s.slot_read;
        d.read_seq_off = s.read_seq_off;
        d.read_seq = s.read_seq;
        d.read_qual = s.read_qual;
        d.read_pos = s.read_pos;
        d.read_end = s.read_end;
        d.read_mapq = s.read_mapq;
        d.read_qcfail = s.read_qcfail;
        d.slot_wi = s.slot_wi;
        d.read_flags = s.read_flags;
        db->shares_reads = true;
    }
    *out = db;
    return PLB_OK;
}

#define CUQ(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return set_err(PLB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                           __LINE__);                                                                      \
    } while (0)

// Plans the tiles of windows [w0, w1), sizes the kernels' shared memory and puts the tile lists on
// the device (async on st).
// zero_copy: the kernels read the tile lists straight from pinned host memory (unified addressing) instead
// of a device copy - a small H2D copy would queue behind the pipelined path's big sequence transfers.
static int plan_chunk(PlbContext* c, PlbDeviceBatch* db, const PlbWindowBatch* hb, int w0, int w1, cudaStream_t st,
                      bool zero_copy = false) {
    db->chunks.emplace_back();
    ChunkPlan& ch = db->chunks.back();
    ch.w0 = w0;
    ch.w1 = w1;
    ch.h0 = hb->win_hap_off[w0];
    ch.h1 = hb->win_hap_off[w1];
    {
        const ByteRanges br = byte_ranges(hb, w0, w1);
        ch.r0 = br.rmax >= br.rmin ? br.rmin : 0;
        ch.r1 = br.rmax >= br.rmin ? br.rmax + 1 : 0;
    }
    int max_read = 0, max_hap = 0, max_H = 0;
    // plan sub-ranges of the chunk on several host threads; their tile lists are concatenated in window order
    // straight into the pinned staging buffer below (no intermediate vector: the lists of a selection call are ~10 MB)
    const int nw = w1 - w0;
    const int parts = std::max(1, std::min(2 * host_threads(), nw / 128));
    std::vector<TileLists> tls(parts);
    std::vector<size_t> off_a((size_t)parts + 1, 0), off_d((size_t)parts + 1, 0);
    {
        std::vector<AnchorPlan> aps(parts);
        std::vector<DpPlan> dps(parts);
        std::vector<int> mr(parts, 0), mh(parts, 0), mH(parts, 0);
        // always the same team size: libgomp rebuilds its thread team whenever num_threads changes
#pragma omp parallel for schedule(static, 1) num_threads(host_threads()) if (parts > 1)
        for (int p = 0; p < parts; ++p) {
            const int a = w0 + (int)((int64_t)nw * p / parts), b = w0 + (int)((int64_t)nw * (p + 1) / parts);
            plan_tiles(hb, a, b, tls[p], aps[p], dps[p], mr[p], mh[p], mH[p]);
        }
        memset(&ch.ap, 0, sizeof ch.ap);
        memset(&ch.dp, 0, sizeof ch.dp);
        for (int p = 0; p < parts; ++p) {
            off_a[(size_t)p + 1] = off_a[(size_t)p] + tls[p].a.size();
            off_d[(size_t)p + 1] = off_d[(size_t)p] + tls[p].d.size();
            AnchorPlan& A = ch.ap;
            const AnchorPlan& q = aps[p];
            A.max_slots = std::max(A.max_slots, q.max_slots);
            A.max_group = std::max(A.max_group, q.max_group);
            A.max_pairs = std::max(A.max_pairs, q.max_pairs);
            A.next_halfs = std::max(A.next_halfs, q.next_halfs);
            A.heads_halfs = std::max(A.heads_halfs, q.heads_halfs);
            A.rpk_words = std::max(A.rpk_words, q.rpk_words);
            A.hpk_words = std::max(A.hpk_words, q.hpk_words);
            DpPlan& D = ch.dp;
            const DpPlan& r = dps[p];
            D.max_slots = std::max(D.max_slots, r.max_slots);
            D.max_group = std::max(D.max_group, r.max_group);
            D.prof_words = std::max(D.prof_words, r.prof_words);
            D.rec_count = std::max(D.rec_count, r.rec_count);
            D.max_pairs = std::max(D.max_pairs, r.max_pairs);
            max_read = std::max(max_read, mr[p]);
            max_hap = std::max(max_hap, mh[p]);
            max_H = std::max(max_H, mH[p]);
        }
        ch.ap.n_tiles = (int)off_a[(size_t)parts];
        ch.dp.n_tiles = (int)off_d[(size_t)parts];
        static const int tma_max = getenv("PLB_DP_TMA_MAX") ? std::min(kTmaMaxLen, atoi(getenv("PLB_DP_TMA_MAX"))) : kTmaMaxLen;
        ch.dp.tma_max = tma_max;
    }
    {
        AnchorPlan& ap = ch.ap;
        ap.max_pairs = (ap.max_pairs + 3) & ~3;
        ap.rpk_words = (ap.rpk_words + 3) & ~3;
        ap.hpk_words = (ap.hpk_words + 3) & ~3;
        ap.next_halfs = (ap.next_halfs + 7) & ~7;
        ap.mult_halfs = 2 * kRankWords;   // the heavy-key bitmap: one bit per 14-bit key
        ap.heads_halfs = std::max(ap.heads_halfs, 4096);
        ap.heads_halfs = (ap.heads_halfs + 7) & ~7;
        ap.cnt_words = (((max_hap + max_read + 2) >> 1) + 3) & ~3;
        const int nwarps = kAnchorThreads / 32;
        // (a vote array for every warp: with two arrays per CTA the rare exact path serialises, 0.98 -> 1.14 ms; a smaller
        // heads area splits the haplotype group more often, 1.14 -> 1.31 ms - both measured while trying to fit six CTAs)
        ap.n_cnt = (int)std::max<size_t>(1, std::min<size_t>(nwarps, kAnchorCntBudget / ((size_t)ap.cnt_words * 4)));
        anchor_layout(ap, sizeof(SlotInfo));
        // ... but one more resident CTA is worth more than the last one or two vote arrays: give up to two of them away
        // when that is what lets another CTA fit (BASELINE config 3: 44.4 KB per tile, four CTAs; 42.8 KB, five)
        auto ctas = [](size_t smem) { return std::min<size_t>(5, (size_t)(220 * 1024) / (smem + 1024)); };
        for (int drop = 1; drop <= 2 && ap.n_cnt - drop >= 6; ++drop) {
            AnchorPlan t = ap;
            t.n_cnt = ap.n_cnt - drop;
            anchor_layout(t, sizeof(SlotInfo));
            if (ctas(t.smem_bytes) > ctas(ap.smem_bytes)) {
                ap = t;
                break;
            }
        }
        ch.a_smem = ap.smem_bytes;
    }
    if (ch.a_smem + 1024 > (size_t)c->smem_optin)
        return set_err(PLB_ERR_SHAPE, "anchor tile needs %zu bytes of shared memory", ch.a_smem);
    // k_anchor exists for 5 and for 4 resident CTAs per SM (48 / 64 registers per thread): launch_windows picks the
    // five-CTA build when this many fit (the run-time modes always take the four-CTA build; a 6 x 40-register build
    // gave the same time as five once the shared memory allowed it, 0.957 vs 0.958 ms)
    ch.a_occ = (int)std::max<size_t>(1, std::min<size_t>(5, (size_t)(220 * 1024) / (ch.a_smem + 1024)));
    // Resident CTAs per SM of the two persistent kernels.  When the chunks of a batch are pipelined over several streams
    // (chunk k+1's anchor kernel next to chunk k's band alignment) the grids are capped so that both fit an SM at once:
    // k_dp 80 registers x 256 threads and ~70 KB of shared memory per CTA, k_anchor 64 x 256 and ~42 KB.
    const int cap_a = getenv("PLB_ANCHOR_OCC") ? std::max(1, atoi(getenv("PLB_ANCHOR_OCC"))) : 0;
    const int cap_d = getenv("PLB_DP_OCC") ? std::max(1, atoi(getenv("PLB_DP_OCC"))) : 0;
    if (cap_a) ch.a_occ = std::min(ch.a_occ, cap_a);
    ch.d_smem = (size_t)ch.dp.prof_words * 4 + (size_t)ch.dp.rec_count * sizeof(HapRec) +
                (size_t)ch.dp.max_slots * (sizeof(DpSlot) + 4) + (size_t)ch.dp.max_group * 4 + (size_t)ch.dp.max_pairs * 12 +
                (size_t)kProfTabWords * 4 + 256 + 64;
    if (ch.d_smem + 1024 > (size_t)c->smem_optin)
        return set_err(PLB_ERR_SHAPE, "dp tile needs %zu bytes of shared memory", ch.d_smem);
    ch.d_occ = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(220 * 1024) / (ch.d_smem + 1024)));
    ch.d_occ = std::min(ch.d_occ, 3);   // __launch_bounds__(256, 3)
    if (cap_d) ch.d_occ = std::min(ch.d_occ, cap_d);
    // Tail splitting: the tiles left over after the last full wave of the persistent DP grid would keep a
    // few CTAs busy for a whole tile time while the others idle.  Cut each of them into k pieces by slot
    // range (k <= 3: a piece still fills the CTA's threads about once) when that shortens the last wave.
    std::vector<Tile> extra;   // the pieces of the split tail tiles
    size_t keep_d = off_d[(size_t)parts];   // DP tiles taken as planned (the rest is replaced by `extra`)
    {
        const int G = c->n_sm * ch.d_occ;
        const int T = (int)off_d[(size_t)parts];
        const int r = T % G;
        if (r > 0 && T > 0) {
            int best_k = 1;
            double best = 1.0;
            for (int k = 2; k <= 3; ++k) {
                const double t = (double)((r * k + G - 1) / G) / k;
                if (t < best - 1e-9) {
                    best = t;
                    best_k = k;
                }
            }
            if (best_k > 1) {
                keep_d = (size_t)(T - r);
                for (int p = 0; p < parts; ++p) {   // the last r tiles, in order
                    const size_t lo = std::max(keep_d, off_d[(size_t)p]), hi = off_d[(size_t)p + 1];
                    for (size_t i = lo; i < hi; ++i) {
                        const Tile& t = tls[p].d[i - off_d[(size_t)p]];
                        const int64_t n = t.s1 - t.s0;
                        const int k = (int)std::min<int64_t>(best_k, std::max<int64_t>(1, n));
                        for (int j = 0; j < k; ++j) {
                            Tile q = t;
                            q.s0 = t.s0 + n * j / k;
                            q.s1 = t.s0 + n * (j + 1) / k;
                            if (q.s1 > q.s0) extra.push_back(q);
                        }
                    }
                }
                ch.dp.n_tiles = (int)(keep_d + extra.size());
            }
        }
    }
    const size_t na = off_a[(size_t)parts] * sizeof(Tile), nd = (keep_d + extra.size()) * sizeof(Tile);
    const size_t na_al = (na + 255) & ~(size_t)255;
    // stage through pinned memory: a copy from a pageable std::vector would block the host until
    // everything queued earlier on the stream (the sequence bytes) has been transferred
    uint8_t* stage = (uint8_t*)pin_alloc(c, na_al + nd + 16);
    if (!stage) return set_err(PLB_ERR_NOMEM, "pinned host allocation failed");
#pragma omp parallel for schedule(static, 1) num_threads(host_threads()) if (parts > 1)
    for (int p = 0; p < parts; ++p) {
        if (!tls[p].a.empty()) memcpy(stage + off_a[(size_t)p] * sizeof(Tile), tls[p].a.data(), tls[p].a.size() * sizeof(Tile));
        const size_t lo = off_d[(size_t)p], hi = std::min(off_d[(size_t)p + 1], keep_d);
        if (hi > lo) memcpy(stage + na_al + lo * sizeof(Tile), tls[p].d.data(), (hi - lo) * sizeof(Tile));
    }
    if (!extra.empty()) memcpy(stage + na_al + keep_d * sizeof(Tile), extra.data(), extra.size() * sizeof(Tile));
    if (zero_copy) {
        ch.ap.tiles = (const Tile*)stage;
        ch.dp.tiles = (const Tile*)(stage + na_al);
        return PLB_OK;
    }
    int rc = block_get(c, na_al + nd + 512, &ch.tiles_blk);
    if (rc) return rc;
    ch.ap.tiles = (const Tile*)ch.tiles_blk.p;
    ch.dp.tiles = (const Tile*)((uint8_t*)ch.tiles_blk.p + na_al);
    if (na) CUQ(cudaMemcpyAsync((void*)ch.ap.tiles, stage, na, cudaMemcpyHostToDevice, st));
    if (nd) CUQ(cudaMemcpyAsync((void*)ch.dp.tiles, stage + na_al, nd, cudaMemcpyHostToDevice, st));
    return PLB_OK;
}

namespace plb {
// slot -> (window, individual) index and haplotype -> window maps, derived on the device
__global__ void k_derive(DevBatch b, int w0, int w1) {
    const int64_t wi0 = (int64_t)w0 * b.n_individuals, wi1 = (int64_t)w1 * b.n_individuals;
    for (int64_t wi = wi0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; wi < wi1; wi += (int64_t)gridDim.x * blockDim.x) {
        int32_t* sw = (int32_t*)b.slot_wi;
        for (int64_t s = b.wi_slot_off[wi]; s < b.wi_slot_off[wi + 1]; ++s) sw[s] = (int32_t)wi;
    }
    for (int w = w0 + blockIdx.x * blockDim.x + threadIdx.x; w < w1; w += gridDim.x * blockDim.x) {
        int32_t* hw = (int32_t*)b.hap_win;
        for (int h = b.win_hap_off[w]; h < b.win_hap_off[w + 1]; ++h) hw[h] = w;
    }
}
}  // namespace plb

// Copies the metadata (everything except the three big byte arrays) that windows [w0, w1) need
// (slot_wi / hap_win are derived by k_derive at the head of the chunk's kernel sequence - never on the
// copy stream, where a kernel would have to wait for SM space behind the persistent kernels and stall
// the DMA queue).  Called once per chunk by the pipelined host path, so the first kernels wait for the
// first chunk's share only.  Per-read arrays follow the same
// "interval already uploaded" rule as the sequence bytes (reads are shared between windows).
static int copy_meta(PlbContext* c, PlbDeviceBatch* db, const PlbWindowBatch* hb, int w0, int w1, int rmin, int rmax,
                     cudaStream_t st) {
    const DevBatch& d = db->d;
    const int nInd = d.n_individuals;
    if (w1 <= w0) return PLB_OK;
    auto cp = [&](const void* dst, const void* src, size_t first, size_t count, size_t esz) -> cudaError_t {
        if (!count || !src) return cudaSuccess;
        return cudaMemcpyAsync((uint8_t*)dst + first * esz, (const uint8_t*)src + first * esz, count * esz,
                               cudaMemcpyHostToDevice, st);
    };
    const size_t nw = (size_t)(w1 - w0);
    const size_t wi0 = (size_t)w0 * nInd, nwi = nw * nInd;
    const size_t h0 = (size_t)hb->win_hap_off[w0], nh = (size_t)hb->win_hap_off[w1] - h0;
    const size_t s0 = (size_t)hb->wi_slot_off[wi0], ns = (size_t)hb->wi_slot_off[wi0 + nwi] - s0;
    CUQ(cp(d.win_hap_off, hb->win_hap_off, w0, nw + 1, 4));
    CUQ(cp(d.hap_seq_off, hb->hap_seq_off, h0, nh + 1, 8));
    CUQ(cp(d.ll_off, db->h_ll_off, wi0, nwi + 1, 8));
    if (db->shares_reads) return PLB_OK;
    CUQ(cp(d.win_start, hb->win_start, w0, nw, 4));
    CUQ(cp(d.win_end, hb->win_end, w0, nw, 4));
    CUQ(cp(d.hap_start, hb->hap_start, w0, nw, 4));
    CUQ(cp(d.wi_slot_off, hb->wi_slot_off, wi0, nwi + 1, 8));
    CUQ(cp(d.wi_n_good, hb->wi_n_good, wi0, nwi, 4));
    CUQ(cp(d.wi_n_bad, hb->wi_n_bad, wi0, nwi, 4));
    CUQ(cp(d.slot_read, hb->slot_read, s0, ns, 4));
    if (db->have_var) {
        CUQ(cp(d.win_n_var, hb->win_n_var, w0, nw, 4));
        CUQ(cp(d.hap_var_mask, hb->hap_var_mask, h0, nh, 8));
        CUQ(cp(d.var_prior, hb->var_prior, (size_t)w0 * hb->max_variants, nw * hb->max_variants, 8));
    }
    // per-read arrays: the parts of [rmin, rmax] not uploaded yet
    if (rmax >= rmin) {
        auto reads = [&](size_t first, size_t count) -> int {
            if (!count) return PLB_OK;
            CUQ(cp(d.read_pos, hb->read_pos, first, count, 4));
            CUQ(cp(d.read_end, hb->read_end, first, count, 4));
            CUQ(cp(d.read_mapq, hb->read_mapq, first, count, 1));
            CUQ(cp(d.read_qcfail, hb->read_qcfail, first, count, 1));
            return PLB_OK;
        };
        int64_t* done = db->readmeta_done;
        const int64_t lo = rmin, hi = (int64_t)rmax + 1;
        int rc;
        if (done[1] <= done[0]) {
            if ((rc = reads((size_t)lo, (size_t)(hi - lo)))) return rc;
            CUQ(cp(d.read_seq_off, hb->read_seq_off, (size_t)lo, (size_t)(hi - lo) + 1, 8));
            done[0] = lo;
            done[1] = hi;
        } else {
            if (lo < done[0]) {
                if ((rc = reads((size_t)lo, (size_t)(done[0] - lo)))) return rc;
                CUQ(cp(d.read_seq_off, hb->read_seq_off, (size_t)lo, (size_t)(done[0] - lo), 8));
                done[0] = lo;
            }
            if (hi > done[1]) {
                if ((rc = reads((size_t)done[1], (size_t)(hi - done[1])))) return rc;
                CUQ(cp(d.read_seq_off, hb->read_seq_off, (size_t)done[1] + 1, (size_t)(hi - done[1]), 8));
                done[1] = hi;
            }
        }
    }
    return PLB_OK;
}

// Copies the parts of [lo, hi) of a per-base array that are not on the device yet; `done` is the interval already
// uploaded (kept as one interval: a gap between intervals is simply filled).  bits = size of an element in the array as
// it travels: 8 (bytes), 2 (packed bases), 4 / 6 (packed qualities); boundary bytes are simply sent again.  The element
// intervals sent are appended to `fresh`.
static int copy_bytes(uint8_t* dst, const uint8_t* src, int64_t lo, int64_t hi, int64_t done[2], cudaStream_t st,
                      int bits = 8, int which = 0, std::vector<PlbDeviceBatch::Fresh>* fresh = nullptr) {
    if (hi <= lo) return PLB_OK;
    auto send = [&](int64_t a, int64_t b) -> cudaError_t {
        if (fresh) fresh->push_back({which, a, b});
        const int64_t b0 = a * bits / 8, b1 = (b * bits + 7) / 8;
        return cudaMemcpyAsync(dst + b0, src + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, st);
    };
    if (done[1] <= done[0]) {
        CUQ(send(lo, hi));
        done[0] = lo;
        done[1] = hi;
        return PLB_OK;
    }
    if (lo < done[0]) {
        CUQ(send(lo, done[0]));
        done[0] = lo;
    }
    if (hi > done[1]) {
        CUQ(send(done[1], hi));
        done[1] = hi;
    }
    return PLB_OK;
}

// Metadata first (small), then the sequence bytes of windows [w0, w1).  `chunk` = which fresh-interval list of the
// batch records what has to be unpacked (packed batches only).
static int copy_seq_for_windows(PlbContext* c, PlbDeviceBatch* db, const PlbWindowBatch* hb, int w0, int w1,
                                cudaStream_t st, int chunk = 0) {
    const ByteRanges r = byte_ranges(hb, w0, w1);
    int rc;
    if ((rc = copy_meta(c, db, hb, w0, w1, r.rmin, r.rmax, st))) return rc;
    int64_t qdone[2] = {db->read_done[0], db->read_done[1]};
    if (db->packed) {
        if (!db->exc_uploaded) {   // the exception lists are small: all of them with the first chunk
            if (db->n_read_exc) {
                CUQ(cudaMemcpyAsync(db->d_read_exc_pos, hb->read_exc_pos, (size_t)db->n_read_exc * 8, cudaMemcpyHostToDevice, st));
                CUQ(cudaMemcpyAsync(db->d_read_exc_chr, hb->read_exc_chr, (size_t)db->n_read_exc, cudaMemcpyHostToDevice, st));
            }
            if (db->n_hap_exc) {
                CUQ(cudaMemcpyAsync(db->d_hap_exc_pos, hb->hap_exc_pos, (size_t)db->n_hap_exc * 8, cudaMemcpyHostToDevice, st));
                CUQ(cudaMemcpyAsync(db->d_hap_exc_chr, hb->hap_exc_chr, (size_t)db->n_hap_exc, cudaMemcpyHostToDevice, st));
            }
            db->exc_uploaded = true;
        }
        auto* fr = &db->fresh[chunk];
        if ((rc = copy_bytes(db->pk_hap, hb->hap_seq, r.hap0, r.hap1, db->hap_done, st, 2, 0, fr))) return rc;
        if ((rc = copy_bytes(db->pk_read, hb->read_seq, r.read0, r.read1, db->read_done, st, 2, 1, fr))) return rc;
    } else {
        if ((rc = copy_bytes((uint8_t*)db->d.hap_seq, hb->hap_seq, r.hap0, r.hap1, db->hap_done, st))) return rc;
        if ((rc = copy_bytes((uint8_t*)db->d.read_seq, hb->read_seq, r.read0, r.read1, db->read_done, st))) return rc;
    }
    if (db->qual_bits) {
        if ((rc = copy_bytes(db->pk_qual, hb->read_qual, r.read0, r.read1, qdone, st, db->qual_bits, 2, &db->fresh[chunk]))) return rc;
    } else {
        if ((rc = copy_bytes((uint8_t*)db->d.read_qual, hb->read_qual, r.read0, r.read1, qdone, st))) return rc;
    }
    return PLB_OK;
}

namespace plb {
// 2-bit packed bases -> the ASCII bytes every kernel reads (A 0, C 1, G 2, T 3; base i at bits 2*(i & 3) of byte i >> 2).
// One thread per packed 32-bit word = sixteen bases = one 16-byte store; writes ONLY bases b0 <= i < b1, so a byte shared
// with a neighbouring interval (unpacked by another chunk, possibly already patched and in use) is never touched twice.
__global__ void __launch_bounds__(256) k_unpack2(const uint8_t* __restrict__ pk, uint8_t* __restrict__ dst, int64_t b0,
                                                 int64_t b1) {
    const int64_t g = (b0 >> 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t base = g << 4;
    if (base >= b1) return;
    const u32 word = ((const u32*)pk)[g];          // the staging area is 256-byte aligned and padded to whole words
    u32 out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const u32 c = (word >> (8 * k)) & 0xFFu;
        const u32 sel = (c & 3u) | ((c & 0xCu) << 2) | ((c & 0x30u) << 4) | ((c & 0xC0u) << 6);
        out[k] = __byte_perm(0x54474341u /* "ACGT" */, 0u, sel);
    }
    if (base >= b0 && base + 16 <= b1) {
        *(uint4*)(dst + base) = make_uint4(out[0], out[1], out[2], out[3]);
    } else {
        for (int k = 0; k < 16; ++k)
            if (base + k >= b0 && base + k < b1) dst[base + k] = (uint8_t)(out[k >> 2] >> (8 * (k & 3)));
    }
}
// Packed qualities -> bytes: code i at bit i * BITS of the stream, value = table[code].  One thread per group of sixteen
// qualities (three 32-bit words at 6 bits, two at 4 bits -> one 16-byte store); the table sits in shared memory (indexing
// the by-value argument with a per-lane index would serialise in the constant cache).  Like k_unpack2 it writes only
// elements b0 <= i < b1.
template <int BITS>
__global__ void __launch_bounds__(256) k_unpack_qual(const uint8_t* __restrict__ pk, uint8_t* __restrict__ dst, int64_t b0,
                                                     int64_t b1, QualTable tab) {
    __shared__ __align__(16) uint8_t s_tab[64];
    if (threadIdx.x < 16) ((u32*)s_tab)[threadIdx.x] = tab.w[threadIdx.x];
    __syncthreads();
    const int64_t g = (b0 >> 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t base = g << 4;
    if (base >= b1) return;
    constexpr int NW = BITS / 2;                                 // 16 * BITS / 32 words per group
    const u32* p = (const u32*)(pk + g * (2 * BITS));            // 16 * BITS / 8 bytes per group: 4-byte aligned
    u32 w[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) w[k] = p[k];
    u32 out[4];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int bit = BITS * k, wi = bit >> 5, sh = bit & 31;
        u32 code = w[wi] >> sh;
        if (sh + BITS > 32) code |= w[wi + 1] << (32 - sh);
        const u32 q = s_tab[code & ((1u << BITS) - 1u)];
        if ((k & 3) == 0) out[k >> 2] = q;
        else out[k >> 2] |= q << (8 * (k & 3));
    }
    if (base >= b0 && base + 16 <= b1) {
        *(uint4*)(dst + base) = make_uint4(out[0], out[1], out[2], out[3]);
    } else {
        for (int k = 0; k < 16; ++k)
            if (base + k >= b0 && base + k < b1) dst[base + k] = (uint8_t)(out[k >> 2] >> (8 * (k & 3)));
    }
}
// ... then the bases that are not A/C/G/T get their original byte back
__global__ void __launch_bounds__(256) k_patch_exceptions(const int64_t* __restrict__ pos, const uint8_t* __restrict__ chr,
                                                          int64_t n, uint8_t* __restrict__ dst, int64_t b0, int64_t b1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t p = pos[i];
    if (p >= b0 && p < b1) dst[p] = chr[i];
}
}  // namespace plb

static int launch_check(PlbContext* c, const char* what);

// Restores the ASCII arrays of the base intervals that `chunk`'s copies brought in (packed batches; no-op otherwise).
static int unpack_fresh(PlbContext* c, PlbDeviceBatch* db, int chunk, cudaStream_t st) {
    if (!db->packed && !db->qual_bits) return PLB_OK;
    int rc;
    for (const auto& f : db->fresh[chunk]) {
        if (f.which == 2) {   // qualities
            const int64_t ng = ((f.hi + 15) >> 4) - (f.lo >> 4);
            if (ng <= 0) continue;
            if (db->qual_bits == 6)
                k_unpack_qual<6><<<(unsigned)((ng + 255) / 256), 256, 0, st>>>(db->pk_qual, (uint8_t*)db->d.read_qual, f.lo, f.hi, db->qtab);
            else
                k_unpack_qual<4><<<(unsigned)((ng + 255) / 256), 256, 0, st>>>(db->pk_qual, (uint8_t*)db->d.read_qual, f.lo, f.hi, db->qtab);
            if ((rc = launch_check(c, "k_unpack_qual"))) return rc;
            continue;
        }
        const uint8_t* pk = f.which ? db->pk_read : db->pk_hap;
        uint8_t* dst = (uint8_t*)(f.which ? db->d.read_seq : db->d.hap_seq);
        const int64_t nq = ((f.hi + 15) >> 4) - (f.lo >> 4);
        if (nq <= 0) continue;
        k_unpack2<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(pk, dst, f.lo, f.hi);
        if ((rc = launch_check(c, "k_unpack2"))) return rc;
        const int64_t ne = f.which ? db->n_read_exc : db->n_hap_exc;
        if (ne > 0) {
            k_patch_exceptions<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(f.which ? db->d_read_exc_pos : db->d_hap_exc_pos,
                                                                             f.which ? db->d_read_exc_chr : db->d_hap_exc_chr, ne,
                                                                             dst, f.lo, f.hi);
            if ((rc = launch_check(c, "k_patch_exceptions"))) return rc;
        }
    }
    db->fresh[chunk].clear();
    return PLB_OK;
}

static int derive_all(PlbContext* c, PlbDeviceBatch* db, cudaStream_t st) {
    const DevBatch& d = db->d;
    if (d.n_windows <= 0) return PLB_OK;
    const int64_t nwi = (int64_t)d.n_windows * d.n_individuals;
    k_derive<<<std::max(1, std::min(4 * c->n_sm, (int)((nwi + 127) / 128))), 128, 0, st>>>(d, 0, d.n_windows);
    return launch_check(c, "k_derive");
}

extern "C" int plb_batch_upload(PlbContext* c, const PlbWindowBatch* hb, PlbDeviceBatch** out) {
    if (!c || !hb || !out) return set_err(PLB_ERR_ARG, "NULL argument");
    int rc = require_idle(c, "plb_batch_upload");
    if (rc) return rc;
    // the kernels index the read pool and the haplotypes unchecked: refuse inconsistent batches here (the scoring options
    // are not known yet; the run-time modes' extra rules are checked by plb_run_device's caller through plb_validate)
    if ((rc = validate_batch(hb, nullptr, 0, false))) return rc;
    PlbDeviceBatch* db = nullptr;
    if ((rc = prepare_batch(c, hb, &db))) return rc;
    cudaStream_t st = c->stream;
    // The device-resident run may cut the batch into chunks that run software-pipelined on the context's streams
    // (PLB_DEVICE_CHUNKS; default one chunk = plain kernel sequence).
    const int dev_chunks = getenv("PLB_DEVICE_CHUNKS") ? std::max(1, std::min(kMaxChunks, atoi(getenv("PLB_DEVICE_CHUNKS")))) : 1;
    const int nch = std::max(1, std::min(dev_chunks, hb->n_windows / kMinChunkWindows));
    db->chunks.reserve(nch);
    rc = copy_seq_for_windows(c, db, hb, 0, hb->n_windows, st);
    if (!rc) rc = unpack_fresh(c, db, 0, st);
    for (int k = 0; k < nch && !rc; ++k) {
        const int wave = c->n_sm * 3;
        auto cutw = [&](int i) {
            int64_t v = (int64_t)hb->n_windows * i / nch;
            if (i > 0 && i < nch && hb->n_windows / nch >= 2 * wave) v = (v + wave / 2) / wave * wave;
            return (int)v;
        };
        if (cutw(k + 1) > cutw(k)) rc = plan_chunk(c, db, hb, cutw(k), cutw(k + 1), st);
    }
    if (rc || (rc = derive_all(c, db, st))) {
        cudaStreamSynchronize(st);
        plb_batch_free(c, db);
        return rc;
    }
    CU(cudaStreamSynchronize(st));
    *out = db;
    return PLB_OK;
}

extern "C" void plb_batch_free(PlbContext* c, PlbDeviceBatch* b) {
    if (!c || !b) return;
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->stream2);
    cudaStreamSynchronize(c->stream3);
    cudaStreamSynchronize(c->copy_stream);
    for (auto& ch : b->chunks) block_put(c, ch.tiles_blk);
    if (b->mode_blk.p) block_put(c, b->mode_blk);
    block_put(c, b->blk);
    delete b;
}

// ---- run --------------------------------------------------------------------------------------

template <typename K>
static int opt_in_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return PLB_OK;
}

static int launch_check(PlbContext* c, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(PLB_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    c->launches++;
    return PLB_OK;
}

// In the run-time modes (flank score, HLA clipping) every band alignment goes through the queue, so
// the default capacity (half the pairs) would overflow into the slow in-kernel fallback: get two
// queues of ~2 entries per pair instead.  Overflow beyond that still falls back, never drops work.
static int mode_queues(PlbContext* c, PlbDeviceBatch* db, bool multi) {
    if (db->mode_blk.p) return PLB_OK;
    const int n = multi ? 3 : 1;
    const int64_t want = std::min<int64_t>((multi ? 1 : 2) * db->d.n_pairs + 4096, (int64_t)1 << 26);
    const size_t qbytes = (((size_t)want * sizeof(QueueEntry)) + 255) & ~(size_t)255;
    int rc = block_get(c, n * (qbytes + 256) + 256, &db->mode_blk);
    if (rc) return rc;
    uint8_t* p = (uint8_t*)db->mode_blk.p;
    Queue* qs[3] = {&db->mq, &db->mq2, &db->mq3};
    for (int i = 0; i < 3; ++i) {
        const int j = i < n ? i : 0;
        qs[i]->e = (QueueEntry*)(p + (size_t)j * (qbytes + 256));
        qs[i]->count = (int32_t*)(p + (size_t)j * (qbytes + 256) + qbytes);
        qs[i]->cap = (int32_t)want;
    }
    return PLB_OK;
}

// Launches the whole kernel sequence for one planned chunk of windows on stream st.  `timed` records
// the per-kernel events of plb_kernel_times (whole-batch launches only).
// `anchor_after` / `anchor_done`: when the chunks of a batch run on several streams they are software-pipelined - chunk
// k+1's anchor kernel starts when chunk k's has finished, i.e. next to chunk k's band-alignment kernel (the two kernels
// are sized to share an SM, see plan_chunk) instead of next to chunk k's anchor kernel.
static int launch_windows(PlbContext* c, PlbDeviceBatch* db, const ChunkPlan& ch, const PlbOptions* opt,
                          PlbPopulationOut* pop, PlbLoglikOut* llo, cudaStream_t st, bool timed, const Queue& q,
                          bool derive = false, int timed_chunk = 0, cudaEvent_t anchor_after = nullptr,
                          cudaEvent_t anchor_done = nullptr) {
    int rc;
    DevBatch& d = db->d;
    const int w0 = ch.w0, w1 = ch.w1;
    if (w1 <= w0) return PLB_OK;
    if (derive) {  // slot -> (window, individual) and haplotype -> window maps of this chunk
        const int64_t nwi = (int64_t)(w1 - w0) * d.n_individuals;
        k_derive<<<std::max(1, std::min(4 * c->n_sm, (int)((nwi + 127) / 128))), 128, 0, st>>>(d, w0, w1);
        if ((rc = launch_check(c, "k_derive"))) return rc;
    }
    ScoreParams sp{opt->gap_extend, opt->nuc_prior, opt->calc_flank_score, opt->use_mapq_cap};
    const int nInd = d.n_individuals;
    const int h0 = ch.h0, h1 = ch.h1;
    CU(cudaMemsetAsync(d.win_flags + w0, 0, (size_t)(w1 - w0) * 4, st));
    CU(cudaMemsetAsync(q.count, 0, 12, st));   // queue fill + the tile counters of k_anchor and k_dp
    const int tslot = c->n_timed % kTimingRing;
    auto mark = [&](int i) {
        if (timed && c->timing) cudaEventRecord(c->kev[tslot][timed_chunk][i], st);
    };
    mark(0);
    if (h1 > h0) {
        k_prep<<<h1 - h0, 128, 0, st>>>(d, h0, sp.ext);
        if ((rc = launch_check(c, "k_prep"))) return rc;
    }
    if (ch.r1 > ch.r0) {   // per-read quality sum / range check (inside the k_prep timing slot)
        k_read_check<<<(unsigned)((ch.r1 - ch.r0 + 255) / 256), 256, 0, st>>>(d, ch.r0, ch.r1, c->d_ctr());
        if ((rc = launch_check(c, "k_read_check"))) return rc;
    }
    mark(1);
    if (ch.ap.n_tiles > 0) {
        const AnchorPlan& ap = ch.ap;
        const bool modes = sp.flank || sp.hla;
        const bool five = !modes && ch.a_occ >= 5;
        const int occ = five ? ch.a_occ : std::min(ch.a_occ, 4);
        if ((rc = five ? opt_in_smem(k_anchor<false, 5>, ch.a_smem)
                       : modes ? opt_in_smem(k_anchor<true, 4>, ch.a_smem) : opt_in_smem(k_anchor<false, 4>, ch.a_smem)))
            return rc;
        const int grid = std::max(1, std::min(ap.n_tiles, c->n_sm * occ));
        if (anchor_after) CU(cudaStreamWaitEvent(st, anchor_after, 0));
        if (modes)
            k_anchor<true, 4><<<grid, kAnchorThreads, ch.a_smem, st>>>(d, ap, q, sp, c->d_ctr());
        else if (five)
            k_anchor<false, 5><<<grid, kAnchorThreads, ch.a_smem, st>>>(d, ap, q, sp, c->d_ctr());
        else
            k_anchor<false, 4><<<grid, kAnchorThreads, ch.a_smem, st>>>(d, ap, q, sp, c->d_ctr());
        if ((rc = launch_check(c, "k_anchor"))) return rc;
        if (anchor_done) CU(cudaEventRecord(anchor_done, st));
        mark(2);
        if (modes)
            k_general<true><<<c->n_sm * 4, 128, 0, st>>>(d, q, sp);
        else
            k_general<false><<<c->n_sm * 4, 128, 0, st>>>(d, q, sp);
        if ((rc = launch_check(c, "k_general"))) return rc;
    } else {
        mark(2);
    }
    mark(3);
    double* ll = (llo && llo->ll) ? llo->ll : db->ll_scratch;
    int32_t* sc = llo ? llo->score : nullptr;
    if (ch.dp.n_tiles > 0) {
        const DpPlan& dp = ch.dp;
        if ((rc = opt_in_smem(k_dp<kDpThreads>, ch.d_smem))) return rc;
        const int grid = std::max(1, std::min(dp.n_tiles, c->n_sm * ch.d_occ));
        k_dp<kDpThreads><<<grid, kDpThreads, ch.d_smem, st>>>(d, dp, sp, ll, sc, q.count + 2, c->d_ctr());
        if ((rc = launch_check(c, "k_dp"))) return rc;
    }
    mark(4);
    if (pop) {
        PopOut po{pop->max_haps, pop->gl,   pop->gl_log_max, pop->gof,       pop->hap_like,
                  pop->freq,     pop->em_post, pop->call,    pop->var_phred, pop->em_iters};
        k_genotype<<<(unsigned)((int64_t)(w1 - w0) * nInd), 64, 0, st>>>(d, ll, po, w0 * nInd);
        if ((rc = launch_check(c, "k_genotype"))) return rc;
        mark(5);
        const int Hm = pop->max_haps;
        const size_t few_smem = (size_t)kPopFewWarps * ((size_t)Hm * (Hm + 1) / 2 + 2 * (size_t)Hm) * 8;
        if (d.n_individuals <= 8 && few_smem <= (size_t)c->smem_optin - 1024) {
            // few individuals: a warp per window, lanes across genotypes (a thread per individual would idle)
            if ((rc = opt_in_smem(k_population_few, few_smem))) return rc;
            k_population_few<<<(w1 - w0 + kPopFewWarps - 1) / kPopFewWarps, 32 * kPopFewWarps, few_smem, st>>>(
                d, po, db->em_scratch, opt->max_em_iters, opt->use_em_likelihoods, w0, w1);
            if ((rc = launch_check(c, "k_population_few"))) return rc;
        } else {
            // thread per individual
            size_t nt = std::min<size_t>(kPopMaxThreads, (size_t)((d.n_individuals + 31) / 32) * 32);
            nt = std::max<size_t>(32, nt / 32 * 32);
            const size_t smem = (size_t)3 * Hm * 8;
            if (smem + 16384 > (size_t)c->smem_optin)
                return set_err(PLB_ERR_SHAPE, "population model needs %zu bytes of shared memory for %d haplotypes", smem, Hm);
            if ((rc = opt_in_smem(k_population, smem))) return rc;
            k_population<<<w1 - w0, (unsigned)nt, smem, st>>>(d, po, db->em_scratch, opt->max_em_iters,
                                                              opt->use_em_likelihoods, (int)nt, w0);
            if ((rc = launch_check(c, "k_population"))) return rc;
        }
    } else {
        mark(5);
    }
    mark(6);
    return PLB_OK;
}

static int check_pop(const PlbDeviceBatch* db, const PlbPopulationOut* pop) {
    if (!pop) return PLB_OK;
    if (pop->max_haps < db->max_haps)
        return set_err(PLB_ERR_SHAPE, "max_haps %d < largest window (%d haplotypes) (cpopulation.pyx:221)", pop->max_haps,
                       db->max_haps);
    if (!pop->gl) return set_err(PLB_ERR_ARG, "PlbPopulationOut.gl is required");
    if (!pop->em_post && pop->max_haps != db->max_haps)
        return set_err(PLB_ERR_ARG, "em_post is NULL: max_haps must equal the batch maximum (%d)", db->max_haps);
    return PLB_OK;
}

extern "C" int plb_run_device(PlbContext* c, PlbDeviceBatch* db, const PlbOptions* opt, PlbPopulationOut* pop,
                              PlbLoglikOut* llo) {
    if (!c || !db) return set_err(PLB_ERR_ARG, "NULL argument");
    int rc = check_options(opt);
    if (rc) return rc;
    if ((rc = require_idle(c, "plb_run_device"))) return rc;
    c->last_done = 0;
    CU(cudaSetDevice(c->device));
    if (db->d.n_windows == 0) return PLB_OK;
    if ((rc = check_pop(db, pop))) return rc;
    cudaStream_t st = c->stream;
    CU(cudaMemsetAsync(c->d_ctr(), 0, sizeof(Counters), st));
    const bool modes = opt->calc_flank_score || opt->use_mapq_cap;
    if (modes && (rc = mode_queues(c, db, db->chunks.size() > 1))) return rc;
    const int nch = (int)db->chunks.size();
    if (nch == 1) {
        if ((rc = launch_windows(c, db, db->chunks[0], opt, pop, llo, st, true, modes ? db->mq : db->q))) return rc;
    } else {
        // chunk k on stream k mod 3, forked from / joined into the context's stream; anchor kernels chained
        if (modes && (rc = mode_queues(c, db, true))) return rc;
        JobSlot& js = c->slot[0];
        cudaStream_t ks[3] = {st, c->stream2, c->stream3};
        const Queue* qs[3] = {modes ? &db->mq : &db->q, modes ? &db->mq2 : &db->q2, modes ? &db->mq3 : &db->q3};
        CU(cudaEventRecord(c->ev_s2, st));
        CU(cudaStreamWaitEvent(c->stream2, c->ev_s2, 0));
        CU(cudaStreamWaitEvent(c->stream3, c->ev_s2, 0));
        for (int k = 0; k < nch; ++k)
            if ((rc = launch_windows(c, db, db->chunks[k], opt, pop, llo, ks[k % 3], true, *qs[k % 3], false, k,
                                     (k > 0 && getenv("PLB_ANCHOR_CHAIN")) ? js.ev_anchor[k - 1] : nullptr, js.ev_anchor[k])))
                return rc;
        CU(cudaEventRecord(c->ev_s2, c->stream2));
        CU(cudaStreamWaitEvent(st, c->ev_s2, 0));
        CU(cudaEventRecord(c->ev_s3, c->stream3));
        CU(cudaStreamWaitEvent(st, c->ev_s3, 0));
    }
    if (c->timing) {
        c->kev_chunks[c->n_timed % kTimingRing] = nch;
        c->n_timed++;
    }
    CU(cudaMemcpyAsync(c->h_ctr(), c->d_ctr(), sizeof(Counters), cudaMemcpyDeviceToHost, st));
    return PLB_OK;
}

static const char* const kErrBitText[] = {
    "a scored read is longer than its haplotype allows (readLen + 15 > hapLen, calign.pyx:256-259)",
    "a base quality above 93",
};

static int counters_status(const Counters* h) {
    if (!h->err) return PLB_OK;
    for (int b = 0; b < 2; ++b)
        if (h->err & (1ull << b)) return set_err(b == 0 ? PLB_ERR_SHAPE : PLB_ERR_ARG, "the kernels refused the batch: %s", kErrBitText[b]);
    return set_err(PLB_ERR_ARG, "the kernels refused the batch (flags %llx)", h->err);
}

extern "C" int plb_last_stats(PlbContext* c, PlbRunStats* out) {
    if (!c || !out) return set_err(PLB_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    const JobSlot& js = c->slot[c->last_done];
    if (!js.busy) CU(cudaStreamSynchronize(c->stream));   // device-resident runs fetch their counters on the stream
    const Counters* h = js.h_ctr;
    out->n_pairs = (int64_t)h->n_pairs;
    out->n_pairs_scored = (int64_t)h->n_scored;
    out->n_dp = (int64_t)h->n_dp;
    out->cells = (int64_t)h->cells;
    out->n_anchor_heavy = (int64_t)h->n_heavy;
    out->n_anchor_verify = (int64_t)h->n_verify;
    out->n_anchor_exact = (int64_t)h->n_exact;
    return counters_status(h);
}

extern "C" int plb_set_timing(PlbContext* c, int on) {
    if (!c) return set_err(PLB_ERR_ARG, "NULL argument");
    c->timing = on != 0;
    c->n_timed = 0;
    return PLB_OK;
}

extern "C" int plb_kernel_times(PlbContext* c, float* ms) {
    if (!c || !ms) return set_err(PLB_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    const int n = c->n_timed < kTimingRing ? c->n_timed : kTimingRing;
    for (int i = 0; i < PLB_N_KERNELS; ++i) {
        double sum = 0.0;
        for (int r = 0; r < n; ++r)
            for (int k = 0; k < c->kev_chunks[r]; ++k) {   // chunks of a pipelined run overlap: their kernel times add up
                float t = 0.f;
                CU(cudaEventElapsedTime(&t, c->kev[r][k][i], c->kev[r][k][i + 1]));
                sum += t;
            }
        ms[i] = n ? (float)(sum / n) : 0.f;
    }
    return n;
}

// ---- host-buffer entry points --------------------------------------------------------------------
//
// Pipelined: the batch is cut into chunks of windows; chunk k+1's sequence bytes travel over PCIe on
// the copy stream while chunk k's kernels run on the compute stream, and every chunk's outputs go
// back as soon as its kernels are done.  With pinned host buffers the call costs about
// max(transfer, compute) instead of their sum.

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Releases a batch whose work is known to be complete (no stream synchronisation: other work may be running).
static void batch_release_done(PlbContext* c, PlbDeviceBatch* b) {
    if (!b) return;
    for (auto& ch : b->chunks) block_put(c, ch.tiles_blk);
    if (b->mode_blk.p) block_put(c, b->mode_blk);
    block_put(c, b->blk);
    delete b;
}

// One host-path call in flight (plb_population_submit .. plb_population_wait).
struct PlbJob {
    int slot = 0;
    PlbDeviceBatch* db = nullptr;
    Block ob{nullptr, 0};
    int n_chunks = 0;
    int rc = PLB_OK;              // error met while queueing (reported by wait, after the queued work has drained)
    char msg[512] = "";
    double t_start = 0, t_prep = 0, t_issue = 0;
    std::vector<cudaEvent_t> tev;     // PLB_TRACE: device timeline
    std::vector<const char*> ttag;
};

static int submit_host(PlbContext* c, const PlbWindowBatch* hb, const PlbOptions* opt, PlbPopulationOut* hpop,
                       PlbLoglikOut* hll, PlbJob** job_out) {
    int rc = check_options(opt);
    if (rc) return rc;
    if (!hb) return set_err(PLB_ERR_ARG, "batch is NULL");
    int slot = -1, busy = 0;
    for (int j = 0; j < kMaxJobs; ++j) {
        if (c->slot[j].busy) ++busy;
        else if (slot < 0) slot = j;
    }
    if (slot < 0) return set_err(PLB_ERR_ARG, "%d jobs already in flight on this context (PLB_MAX_JOBS); wait for one first", kMaxJobs);
    CU(cudaSetDevice(c->device));
    // O(slots) consistency check (the kernels index the pool unchecked); base qualities are checked on the GPU
    if ((rc = validate_batch(hb, opt, hpop ? hpop->max_haps : 0, false))) return rc;
    PlbJob* job = new PlbJob();
    job->slot = slot;
    *job_out = job;
    if (hb->n_windows == 0) return PLB_OK;   // nothing queued; wait() returns at once
    c->cur = slot;
    JobSlot& js = c->slot[slot];
    static const bool trace = getenv("PLB_TRACE") != nullptr;
    job->t_start = now_ms();
    PlbDeviceBatch* db = nullptr;
    if ((rc = prepare_batch(c, hb, &db))) {
        delete job;
        *job_out = nullptr;
        return rc;
    }
    job->db = db;
    job->t_prep = now_ms();
    const int W = hb->n_windows, nInd = hb->n_individuals;
    const int64_t n_pairs = db->d.n_pairs;
    Block& ob = job->ob;
    PlbPopulationOut dpop;
    PlbLoglikOut dll;
    memset(&dpop, 0, sizeof dpop);
    memset(&dll, 0, sizeof dll);
    Layout L;
    size_t o_gl = 0, o_glmax = 0, o_gof = 0, o_hl = 0, o_freq = 0, o_em = 0, o_call = 0, o_vp = 0, o_it = 0, o_ll = 0,
           o_sc = 0;
    int Hm = 0, Gm = 0, V = hb->max_variants;
    auto abandon = [&](int code) {
        batch_release_done(c, db);
        if (ob.p) block_put(c, ob);
        delete job;
        *job_out = nullptr;
        return code;
    };
    if (hpop) {
        Hm = hpop->max_haps;
        if (Hm < db->max_haps)
            return abandon(set_err(PLB_ERR_SHAPE, "max_haps %d < largest window (%d haplotypes) (cpopulation.pyx:221)", Hm,
                                   db->max_haps));
        Gm = Hm * (Hm + 1) / 2;
        o_gl = L.take((size_t)W * nInd * Gm * 8);
        o_glmax = L.take((size_t)W * nInd * 8);
        o_gof = L.take(hpop->gof ? (size_t)W * Gm * nInd * 8 : 0);
        o_hl = L.take(hpop->hap_like ? (size_t)W * nInd * Hm * 8 : 0);
        o_freq = L.take((size_t)W * Hm * 8);
        o_em = L.take((size_t)W * nInd * Gm * 8);
        o_call = L.take((size_t)W * nInd * 4);
        o_vp = L.take((size_t)W * std::max(V, 1) * 8);
        o_it = L.take((size_t)W * 4);
    }
    const bool want_ll = hll && hll->ll, want_sc = hll && hll->score;
    if (want_ll) o_ll = L.take((size_t)n_pairs * 8);
    if (want_sc) o_sc = L.take((size_t)n_pairs * 4);
    if ((rc = block_get(c, L.off + 256, &ob))) return abandon(rc);
    if (hpop) {
        dpop.max_haps = Hm;
        dpop.gl = at<double>(ob, o_gl);
        dpop.gl_log_max = at<double>(ob, o_glmax);
        dpop.gof = hpop->gof ? at<double>(ob, o_gof) : nullptr;
        dpop.hap_like = hpop->hap_like ? at<double>(ob, o_hl) : nullptr;
        dpop.freq = at<double>(ob, o_freq);
        dpop.em_post = at<double>(ob, o_em);
        dpop.call = at<int32_t>(ob, o_call);
        dpop.var_phred = (V > 0 && db->have_var && hpop->var_phred) ? at<double>(ob, o_vp) : nullptr;
        dpop.em_iters = at<int32_t>(ob, o_it);
    }
    dll.ll = want_ll ? at<double>(ob, o_ll) : nullptr;
    dll.score = want_sc ? at<int32_t>(ob, o_sc) : nullptr;

    cudaStream_t cs = c->copy_stream, st = c->stream;
    cudaStream_t kst = st;   // compute stream of the current chunk (chunks rotate over three streams so that one
                             // chunk's kernel tails overlap the next chunk's kernels)
    cudaError_t e = cudaSuccess;
    auto d2h = [&](void* dst, const void* src, size_t off, size_t bytes) {
        if (rc == PLB_OK && e == cudaSuccess && dst && src && bytes)
            e = cudaMemcpyAsync((uint8_t*)dst + off, (const uint8_t*)src + off, bytes, cudaMemcpyDefault, c->aux_stream);   // dst: host or device
    };
    // Chunking.  A lone call wants its first kernels early: up to six chunks whose sizes grow x1.3.  When another job
    // is still computing, this job's bytes travel behind that job's kernels anyway, so it is cut into three equal chunks
    // on the three compute streams (one chunk's kernel tails and unpacking run under the next chunk's kernels; measured
    // 4.58 / 4.51 / 4.46 ms per step with 1 / 2 / 3 chunks); PLB_PIPE_CHUNKS overrides that number.
    const int pipe_chunks = getenv("PLB_PIPE_CHUNKS") ? std::max(1, std::min(kMaxChunks, atoi(getenv("PLB_PIPE_CHUNKS")))) : 3;
    int n_chunks = std::max(1, std::min(kMaxChunks, W / kMinChunkWindows));
    const bool pipelined = busy > 0;
    if (pipelined) n_chunks = std::min(n_chunks, pipe_chunks);
    job->n_chunks = n_chunks;
    db->chunks.reserve(n_chunks);
    const bool modes = opt->calc_flank_score || opt->use_mapq_cap;
    rc = modes ? mode_queues(c, db, n_chunks > 1) : PLB_OK;
    // PLB_TRACE: device timeline of the pipeline (events with timing, created per call)
    auto tmark_t = [&](cudaStream_t s_, const char* tag) {
        if (!trace) return;
        cudaEvent_t ev_;
        cudaEventCreate(&ev_);
        cudaEventRecord(ev_, s_);
        job->tev.push_back(ev_);
        job->ttag.push_back(tag);
    };
    auto tmark = [&](cudaStream_t s_) { tmark_t(s_, s_ == cs ? "h2d:" : "k:"); };
    tmark(cs);
    if (rc == PLB_OK) e = cudaMemsetAsync(js.d_ctr, 0, sizeof(Counters), cs);
    cudaStream_t kstreams[3] = {st, c->stream2, c->stream3};
    const Queue* kqueues[3] = {modes ? &db->mq : &db->q, modes ? &db->mq2 : &db->q2, modes ? &db->mq3 : &db->q3};
    // Chunk sizes grow geometrically (x1.3): the first kernels start after a small upload, and because
    // the kernels need ~1.3x the time of the PCIe transfer of the same windows, chunk k+1 has just
    // arrived when chunk k finishes.  Sizes are rounded to whole waves of the DP grid (one DP tile
    // per window in the common case) so that no chunk ends in a nearly empty wave.
    auto cut = [&](int i) -> int {
        if (i <= 0) return 0;
        if (i >= n_chunks) return W;
        const double r = pipelined ? 1.0 : 1.3;
        double tot = 0.0, part = 0.0, term = 1.0;
        for (int j = 0; j < n_chunks; ++j, term *= r) {
            tot += term;
            if (j < i) part += term;
        }
        int64_t v = (int64_t)((double)W * part / tot);
        const int64_t wave = (int64_t)c->n_sm * 3;
        if (W / n_chunks >= 2 * wave) v = std::max<int64_t>(wave, (v + wave / 2) / wave * wave);
        return (int)std::min<int64_t>(v, W);
    };
    // The DMA must never wait for the host: the copies of chunk k+1 are queued BEFORE chunk k is planned,
    // and once the first chunk is launched all remaining copies are queued at once.
    const bool chain_anchors = getenv("PLB_ANCHOR_CHAIN") != nullptr;   // measured: no gain (profiles/sweep_overlap_r02.txt)
    int copies_queued = 0;
    auto queue_copies = [&](int upto) {
        for (; copies_queued < upto && copies_queued < n_chunks && rc == PLB_OK && e == cudaSuccess; ++copies_queued) {
            const int a = cut(copies_queued), b2 = cut(copies_queued + 1);
            if (b2 > a) rc = copy_seq_for_windows(c, db, hb, a, b2, cs, 1 + copies_queued);
            if (rc == PLB_OK) e = cudaEventRecord(js.ev_chunk[1 + copies_queued], cs);
            tmark(cs);
        }
    };
    for (int k = 0; k < n_chunks && rc == PLB_OK && e == cudaSuccess; ++k) {
        // lone job: the DMA must not wait for the host (two chunks queued ahead, then all of them); queued job: the DMA
        // still works on the previous job's bytes, so chunk k's bytes and tile lists are queued in turn and chunk k's
        // kernels need nothing that travels behind chunk k+1
        queue_copies(pipelined ? k + 1 : (k == 0 ? 2 : n_chunks));
        if (rc != PLB_OK || e != cudaSuccess) break;
        const int w0 = cut(k), w1 = cut(k + 1);
        if (w1 <= w0) continue;
        kst = kstreams[k % 3];
        // A lone job's kernels read their tile lists from pinned host memory (a small H2D copy would wait behind the
        // sequence bytes).  A job queued behind another one has time: its lists follow its bytes on the copy stream, so its
        // persistent kernels fetch tiles from HBM instead of across a PCIe link that the next job's upload keeps busy.
        // (For a lone job, sending the lists on another stream was measured: the copy engine serves them behind the queued
        // sequence bytes all the same, 6.2 -> 9.6 ms per call.)
        if ((rc = plan_chunk(c, db, hb, w0, w1, cs, !pipelined))) break;
        if (pipelined) {
            e = cudaEventRecord(js.ev_tiles[k], cs);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(kst, js.ev_tiles[k], 0);
            if (e != cudaSuccess) break;
        }
        e = cudaStreamWaitEvent(kst, js.ev_chunk[1 + k], 0);
        if (e != cudaSuccess) break;
        tmark_t(kst, "u:");
        if (db->packed || db->qual_bits) {
            // reads are shared between chunks: chunk k may score reads that chunk k-1's copies brought in, so its
            // kernels also wait for that chunk's unpacking (which waited for the one before it)
            if (k > 0) e = cudaStreamWaitEvent(kst, js.ev_unpack[k - 1], 0);
            if (e != cudaSuccess) break;
            if ((rc = unpack_fresh(c, db, 1 + k, kst))) break;
            e = cudaEventRecord(js.ev_unpack[k], kst);
            if (e != cudaSuccess) break;
        }
        tmark(kst);
        if ((rc = launch_windows(c, db, db->chunks.back(), opt, hpop ? &dpop : nullptr, &dll, kst, false,
                                 *kqueues[k % 3], true, 0, (chain_anchors && k > 0) ? js.ev_anchor[k - 1] : nullptr,
                                 js.ev_anchor[k])))
            break;
        tmark(kst);
        // the chunk's outputs go back on the D2H stream, behind an event: the compute stream is free for the next kernels
        if (e == cudaSuccess) e = cudaEventRecord(js.ev_kdone[k], kst);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->aux_stream, js.ev_kdone[k], 0);
        if (hpop) {
            const size_t wq = (size_t)w0, wn = (size_t)(w1 - w0);
            d2h(hpop->gl, dpop.gl, wq * nInd * Gm * 8, wn * nInd * Gm * 8);
            d2h(hpop->gl_log_max, dpop.gl_log_max, wq * nInd * 8, wn * nInd * 8);
            d2h(hpop->gof, dpop.gof, wq * Gm * nInd * 8, wn * Gm * nInd * 8);
            d2h(hpop->hap_like, dpop.hap_like, wq * nInd * Hm * 8, wn * nInd * Hm * 8);
            d2h(hpop->freq, dpop.freq, wq * Hm * 8, wn * Hm * 8);
            d2h(hpop->em_post, dpop.em_post, wq * nInd * Gm * 8, wn * nInd * Gm * 8);
            d2h(hpop->call, dpop.call, wq * nInd * 4, wn * nInd * 4);
            if (V > 0) d2h(hpop->var_phred, dpop.var_phred, wq * V * 8, wn * V * 8);
            d2h(hpop->em_iters, dpop.em_iters, wq * 4, wn * 4);
        }
        const int64_t p0 = db->h_ll_off[(size_t)w0 * nInd], p1 = db->h_ll_off[(size_t)w1 * nInd];
        if (want_ll) d2h(hll->ll, dll.ll, (size_t)p0 * 8, (size_t)(p1 - p0) * 8);
        if (want_sc) d2h(hll->score, dll.score, (size_t)p0 * 4, (size_t)(p1 - p0) * 4);
        tmark_t(c->aux_stream, "d2h:");
    }
    // completion: one event per compute stream (no join through the first stream - the next job's first chunk must not
    // wait for this job's last chunks on the other two), the counters come back on the side stream once all three fired
    for (int i = 0; i < 3; ++i) {
        if (e == cudaSuccess) e = cudaEventRecord(js.ev_done[i], kstreams[i]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->aux_stream, js.ev_done[i], 0);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(js.h_ctr, js.d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, c->aux_stream);
    if (e == cudaSuccess) e = cudaEventRecord(js.ev_ctr, c->aux_stream);
    job->t_issue = now_ms();
    if (rc == PLB_OK && e != cudaSuccess) rc = set_err(PLB_ERR_CUDA, "pipelined run failed: %s", cudaGetErrorString(e));
    job->rc = rc;
    if (rc != PLB_OK) snprintf(job->msg, sizeof job->msg, "%s", g_err);
    js.busy = true;
    c->jobs_submitted++;
    return PLB_OK;   // queueing errors are reported by wait, after what was queued has drained
}

static int wait_job(PlbContext* c, PlbJob* job) {
    if (!job) return set_err(PLB_ERR_ARG, "job is NULL");
    static const bool trace = getenv("PLB_TRACE") != nullptr;
    int rc = job->rc;
    if (job->db) {
        JobSlot& js = c->slot[job->slot];
        cudaSetDevice(c->device);
        const double t_w0 = now_ms();
        // the copy stream only carries this job's uploads up to ev_chunk[n]; the compute side ends with ev_ctr
        cudaError_t e2 = cudaEventSynchronize(js.ev_chunk[job->n_chunks]);
        const double t_copy = now_ms();
        cudaError_t e3 = cudaEventSynchronize(js.ev_ctr);
        for (int i = 0; i < 3 && e3 == cudaSuccess; ++i) e3 = cudaEventSynchronize(js.ev_done[i]);
        if (rc != PLB_OK) {   // a queueing error: drain everything so that nothing still reads the caller's buffers
            cudaStreamSynchronize(c->copy_stream);
            cudaStreamSynchronize(c->stream);
            cudaStreamSynchronize(c->stream2);
            cudaStreamSynchronize(c->stream3);
            cudaStreamSynchronize(c->aux_stream);
        }
        if (trace && !job->tev.empty()) {
            fprintf(stderr, "[plb] timeline (ms):");
            for (size_t i = 1; i < job->tev.size(); ++i) {
                float t = 0;
                cudaEventElapsedTime(&t, job->tev[0], job->tev[i]);
                fprintf(stderr, " %s%.2f", job->ttag[i], t);
            }
            // idle time of the compute side between jobs: from the previous job's last compute mark to this job's first
            int first_k = -1, last_k = -1;
            for (size_t i = 1; i < job->tev.size(); ++i)
                if (job->ttag[i][0] == 'u' || job->ttag[i][0] == 'k') {
                    if (first_k < 0) first_k = (int)i;
                    last_k = (int)i;
                }
            if (c->trace_prev_end && first_k >= 0) {
                float t = 0;
                if (cudaEventElapsedTime(&t, c->trace_prev_end, job->tev[first_k]) == cudaSuccess)
                    fprintf(stderr, "  [previous job's last kernel mark -> this job's first: %+.2f ms]", t);
            }
            if (last_k >= 0) {
                if (c->trace_prev_end) cudaEventDestroy(c->trace_prev_end);
                c->trace_prev_end = job->tev[last_k];
                job->tev[last_k] = nullptr;
            }
            fprintf(stderr, "\n");
        }
        for (cudaEvent_t ev_ : job->tev)
            if (ev_) cudaEventDestroy(ev_);
        if (trace)
            fprintf(stderr, "[plb] job slot %d: prepare %.2f ms, issue %.2f ms, submit->wait %.2f ms, copy drain +%.2f ms, compute drain +%.2f ms (%d chunks)\n",
                    job->slot, job->t_prep - job->t_start, job->t_issue - job->t_prep, t_w0 - job->t_issue, t_copy - t_w0,
                    now_ms() - t_copy, job->n_chunks);
        if (rc == PLB_OK && e2 != cudaSuccess) rc = set_err(PLB_ERR_CUDA, "copy stream: %s", cudaGetErrorString(e2));
        if (rc == PLB_OK && e3 != cudaSuccess) rc = set_err(PLB_ERR_CUDA, "compute stream: %s", cudaGetErrorString(e3));
        else if (rc != PLB_OK && job->msg[0]) set_err(rc, "%s", job->msg);
        if (rc == PLB_OK) rc = counters_status(js.h_ctr);
        block_put(c, job->ob);
        batch_release_done(c, job->db);
        js.busy = false;
        c->last_done = job->slot;
    }
    delete job;
    return rc;
}

extern "C" int plb_population_submit(PlbContext* c, const PlbWindowBatch* hb, const PlbOptions* opt, PlbPopulationOut* out,
                                     PlbLoglikOut* ll, PlbJob** job) {
    if (!c || !job) return set_err(PLB_ERR_ARG, "NULL argument");
    *job = nullptr;
    if (!out && !(ll && (ll->ll || ll->score))) return set_err(PLB_ERR_ARG, "no output requested");
    if (out && !out->gl) return set_err(PLB_ERR_ARG, "PlbPopulationOut.gl is required");
    return submit_host(c, hb, opt, out, ll, job);
}

extern "C" int plb_population_wait(PlbContext* c, PlbJob* job) {
    if (!c) return set_err(PLB_ERR_ARG, "NULL argument");
    return wait_job(c, job);
}

static int run_host(PlbContext* c, const PlbWindowBatch* hb, const PlbOptions* opt, PlbPopulationOut* hpop,
                    PlbLoglikOut* hll) {
    PlbJob* job = nullptr;
    int rc = submit_host(c, hb, opt, hpop, hll, &job);
    if (rc) return rc;
    return wait_job(c, job);
}

extern "C" int plb_window_loglik_host(PlbContext* c, const PlbWindowBatch* hb, const PlbOptions* opt,
                                      PlbLoglikOut* out) {
    if (!c || !out) return set_err(PLB_ERR_ARG, "NULL argument");
    return run_host(c, hb, opt, nullptr, out);
}

extern "C" int plb_population_run_host(PlbContext* c, const PlbWindowBatch* hb, const PlbOptions* opt,
                                       PlbPopulationOut* out, PlbLoglikOut* ll) {
    if (!c || !out) return set_err(PLB_ERR_ARG, "NULL argument");
    if (!out->gl) return set_err(PLB_ERR_ARG, "PlbPopulationOut.gl is required");
    return run_host(c, hb, opt, out, ll);
}

// ---- packing helpers (host) ------------------------------------------------------------------------

namespace {
// code of a byte: 0..3 for exactly 'A','C','G','T', 4 otherwise
inline int acgt_code(uint8_t ch) { return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 4; }

// Shared body: base i of the source is get(i) (an ASCII byte).  Bits are OR-ed into dst, so a partial first byte that
// already holds earlier bases is preserved.  Big arrays are cut at byte boundaries of dst and packed in parallel; the
// exceptions of the pieces are concatenated in order.
template <typename Get>
int pack_2bit(Get get, int64_t n, uint8_t* dst, int64_t dst_base, int64_t* exc_pos, uint8_t* exc_chr, int64_t exc_cap,
              int64_t* n_exc) {
    if (n < 0 || dst_base < 0 || !dst || !n_exc || *n_exc < 0) return set_err(PLB_ERR_ARG, "NULL / bad argument");
    const int parts = n >= (1 << 20) ? host_threads() : 1;
    std::vector<std::vector<std::pair<int64_t, uint8_t>>> exc((size_t)parts);
#pragma omp parallel for schedule(static, 1) num_threads(host_threads()) if (parts > 1)
    for (int p = 0; p < parts; ++p) {
        // piece boundaries on multiples of 4 of the destination index
        auto bound = [&](int q) -> int64_t {
            if (q <= 0) return 0;
            if (q >= parts) return n;
            const int64_t t = (dst_base + n * q / parts) & ~(int64_t)3;
            return std::min(n, std::max<int64_t>(0, t - dst_base));
        };
        const int64_t i0 = bound(p), i1 = bound(p + 1);
        for (int64_t i = i0; i < i1; ++i) {
            const uint8_t ch = get(i);
            int code = acgt_code(ch);
            const int64_t j = dst_base + i;
            if (code > 3) {
                exc[(size_t)p].push_back({j, ch});
                code = 0;
            }
            dst[j >> 2] |= (uint8_t)(code << (2 * (j & 3)));
        }
    }
    int64_t k = *n_exc;
    for (auto& v : exc)
        for (auto& e : v) {
            if (k >= exc_cap || !exc_pos || !exc_chr)
                return set_err(PLB_ERR_SHAPE, "more than %lld bases outside ACGT: exception arrays too small", (long long)exc_cap);
            exc_pos[k] = e.first;
            exc_chr[k] = e.second;
            ++k;
        }
    *n_exc = k;
    return PLB_OK;
}
}  // namespace

extern "C" int plb_pack_quals_host(const uint8_t* src, int64_t n, uint8_t* dst, int32_t* qual_bits, uint8_t* qual_table) {
    if (n < 0 || (n > 0 && !src) || !dst || !qual_bits || !qual_table) return set_err(PLB_ERR_ARG, "NULL / bad argument");
    // distinct values (a histogram per thread, then merged)
    const int nt = n >= (1 << 20) ? host_threads() : 1;
    std::vector<uint8_t> seen((size_t)nt * 256, 0);
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (nt > 1)
    for (int t = 0; t < nt; ++t) {
        uint8_t* sn = seen.data() + (size_t)t * 256;
        const int64_t a = n * t / nt, b = n * (t + 1) / nt;
        for (int64_t i = a; i < b; ++i) sn[src[i]] = 1;
    }
    uint8_t code[256];
    int nd = 0;
    for (int v = 0; v < 256; ++v) {
        bool any = false;
        for (int t = 0; t < nt; ++t) any |= seen[(size_t)t * 256 + v] != 0;
        if (!any) continue;
        if (v > 93) return set_err(PLB_ERR_SHAPE, "base quality %d > 93", v);
        if (nd == 64) return set_err(PLB_ERR_SHAPE, "more than 64 distinct base qualities: the batch stays at 8 bits");
        code[v] = (uint8_t)nd;
        qual_table[nd++] = (uint8_t)v;
    }
    for (int i = nd; i < 64; ++i) qual_table[i] = 0;
    const int bits = nd <= 16 ? 4 : 6;
    *qual_bits = bits;
    const int64_t ng = (n + 3) / 4;   // groups of four qualities = bits / 2 bytes each
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (ng > (1 << 18))
    for (int64_t g = 0; g < ng; ++g) {
        uint32_t v = 0;
        for (int k = 0; k < 4; ++k) {
            const int64_t i = 4 * g + k;
            if (i < n) v |= (uint32_t)code[src[i]] << (bits * k);
        }
        uint8_t* p = dst + g * (bits / 2);
        p[0] = (uint8_t)v;
        p[1] = (uint8_t)(v >> 8);
        if (bits == 6) p[2] = (uint8_t)(v >> 16);
    }
    return PLB_OK;
}

extern "C" int plb_pack_bases_host(const uint8_t* src, int64_t n, uint8_t* dst, int64_t dst_base, int64_t* exc_pos,
                                   uint8_t* exc_chr, int64_t exc_cap, int64_t* n_exc) {
    if (n > 0 && !src) return set_err(PLB_ERR_ARG, "src is NULL");
    return pack_2bit([&](int64_t i) { return src[i]; }, n, dst, dst_base, exc_pos, exc_chr, exc_cap, n_exc);
}

extern "C" int plb_pack_nibbles_host(const uint8_t* bam_seq, int64_t n, uint8_t* dst, int64_t dst_base, int64_t* exc_pos,
                                     uint8_t* exc_chr, int64_t exc_cap, int64_t* n_exc) {
    if (n > 0 && !bam_seq) return set_err(PLB_ERR_ARG, "bam_seq is NULL");
    static const char* const kNib = "=ACMGRSVTWYHKDBN";   // htslibWrapper.pyx:414-416
    return pack_2bit([&](int64_t i) { return (uint8_t)kNib[(bam_seq[i >> 1] >> (4 * (1 - (i & 1)))) & 15]; }, n, dst, dst_base,
                     exc_pos, exc_chr, exc_cap, n_exc);
}

// ---- N4: per-site genotype calls ------------------------------------------------------------------

extern "C" int plb_site_genotypes_host(PlbContext* c, const PlbWindowBatch* hb, const PlbPopulationOut* pop,
                                       const PlbSiteBatch* st, PlbSiteOut* out) {
    if (!c || !hb || !pop || !st || !out) return set_err(PLB_ERR_ARG, "NULL argument");
    if (!pop->gl || !pop->gof || !pop->freq) return set_err(PLB_ERR_ARG, "PlbPopulationOut needs gl, gof and freq");
    if (!hb->win_hap_off || !hb->hap_var_mask || !hb->wi_n_good)
        return set_err(PLB_ERR_ARG, "batch needs win_hap_off, hap_var_mask and wi_n_good");
    const int S = st->n_sites, W = hb->n_windows, nInd = hb->n_individuals;
    if (S < 0 || nInd < 1) return set_err(PLB_ERR_ARG, "bad n_sites / n_individuals");
    if (S == 0) return PLB_OK;
    if (!st->site_win || !st->site_var_off || !st->site_hap_off || !st->hap_is_ref)
        return set_err(PLB_ERR_ARG, "NULL array in site batch");
    const int Hm = pop->max_haps, Gm = Hm * (Hm + 1) / 2, P = out->max_pairs;
    for (int s = 0; s < S; ++s) {
        const int w = st->site_win[s];
        if (w < 0 || w >= W) return set_err(PLB_ERR_ARG, "site %d: window %d out of range", s, w);
        const int H = hb->win_hap_off[w + 1] - hb->win_hap_off[w];
        if (H > Hm) return set_err(PLB_ERR_SHAPE, "site %d: window has %d haplotypes > max_haps %d", s, H, Hm);
        if (st->site_hap_off[s + 1] - st->site_hap_off[s] != H)
            return set_err(PLB_ERR_ARG, "site %d: hap_is_ref must hold one entry per haplotype of its window", s);
        const int nV = st->site_var_off[s + 1] - st->site_var_off[s];
        if (nV < 0 || (nV + 1) * (nV + 2) / 2 > P)
            return set_err(PLB_ERR_SHAPE, "site %d: %d variants need %d allele pairs > max_pairs %d", s, nV,
                           (nV + 1) * (nV + 2) / 2, P);
        for (int k = 0; k < nV; ++k) {
            const int v = st->site_var[st->site_var_off[s] + k];
            if (v < 0 || v >= 64) return set_err(PLB_ERR_ARG, "site %d: variant index %d out of range", s, v);
        }
    }
    CU(cudaSetDevice(c->device));
    const int n_haps = hb->win_hap_off[W];
    const int64_t n_sv = st->site_var_off[S], n_sh = st->site_hap_off[S];
    const size_t SI = (size_t)S * nInd;
    Layout L;
    const size_t o_sw = L.take((size_t)S * 4), o_svo = L.take((size_t)(S + 1) * 4), o_sv = L.take((size_t)n_sv * 4 + 4),
                 o_sho = L.take((size_t)(S + 1) * 8), o_ref = L.take((size_t)n_sh + 4), o_who = L.take((size_t)(W + 1) * 4),
                 o_mask = L.take((size_t)n_haps * 8), o_ng = L.take((size_t)W * nInd * 4),
                 o_gl = L.take((size_t)W * nInd * Gm * 8), o_gof = L.take((size_t)W * Gm * nInd * 8),
                 o_fr = L.take((size_t)W * Hm * 8);
    const size_t o_ph = L.take(out->phased ? SI * 8 : 0), o_lik = L.take(out->lik ? SI * P * 8 : 0),
                 o_post = L.take(out->post ? SI * 24 : 0), o_phr = L.take(out->phred ? SI * 12 : 0),
                 o_og = L.take(out->gof ? SI * 8 : 0), o_gt = L.take(out->gt ? SI * 8 : 0),
                 o_l10 = L.take(out->gl_log10 ? SI * 24 : 0);
    Block B;
    int rc = block_get(c, L.off + 256, &B);
    if (rc) return rc;
    cudaStream_t stq = c->stream;
    cudaError_t e = cudaSuccess;
    auto up = [&](size_t off, const void* src, size_t bytes) {
        if (e == cudaSuccess && bytes) e = cudaMemcpyAsync((uint8_t*)B.p + off, src, bytes, cudaMemcpyHostToDevice, stq);
    };
    up(o_sw, st->site_win, (size_t)S * 4);
    up(o_svo, st->site_var_off, (size_t)(S + 1) * 4);
    up(o_sv, st->site_var, (size_t)n_sv * 4);
    up(o_sho, st->site_hap_off, (size_t)(S + 1) * 8);
    up(o_ref, st->hap_is_ref, (size_t)n_sh);
    up(o_who, hb->win_hap_off, (size_t)(W + 1) * 4);
    up(o_mask, hb->hap_var_mask, (size_t)n_haps * 8);
    up(o_ng, hb->wi_n_good, (size_t)W * nInd * 4);
    up(o_gl, pop->gl, (size_t)W * nInd * Gm * 8);
    up(o_gof, pop->gof, (size_t)W * Gm * nInd * 8);
    up(o_fr, pop->freq, (size_t)W * Hm * 8);
    if (e == cudaSuccess) {
        SiteIn in{S, nInd, Hm, st->min_posterior, at<int32_t>(B, o_sw), at<int32_t>(B, o_svo), at<int32_t>(B, o_sv),
                  at<int64_t>(B, o_sho), at<uint8_t>(B, o_ref), at<int32_t>(B, o_who), at<uint64_t>(B, o_mask),
                  at<int32_t>(B, o_ng), at<double>(B, o_gl), at<double>(B, o_gof), at<double>(B, o_fr)};
        SiteOutDev od{P,
                      out->phased ? at<int32_t>(B, o_ph) : nullptr,
                      out->lik ? at<double>(B, o_lik) : nullptr,
                      out->post ? at<double>(B, o_post) : nullptr,
                      out->phred ? at<int32_t>(B, o_phr) : nullptr,
                      out->gof ? at<double>(B, o_og) : nullptr,
                      out->gt ? at<int32_t>(B, o_gt) : nullptr,
                      out->gl_log10 ? at<double>(B, o_l10) : nullptr};
        k_site_genotypes<<<(unsigned)((SI + 127) / 128), 128, 0, stq>>>(in, od);
        e = cudaGetLastError();
        c->launches++;
    }
    auto down = [&](void* dst, size_t off, size_t bytes) {
        if (e == cudaSuccess && dst && bytes) e = cudaMemcpyAsync(dst, (uint8_t*)B.p + off, bytes, cudaMemcpyDeviceToHost, stq);
    };
    down(out->phased, o_ph, SI * 8);
    down(out->lik, o_lik, SI * P * 8);
    down(out->post, o_post, SI * 24);
    down(out->phred, o_phr, SI * 12);
    down(out->gof, o_og, SI * 8);
    down(out->gt, o_gt, SI * 8);
    down(out->gl_log10, o_l10, SI * 24);
    cudaError_t e2 = cudaStreamSynchronize(stq);
    block_put(c, B);
    if (e != cudaSuccess) return set_err(PLB_ERR_CUDA, "plb_site_genotypes_host: %s", cudaGetErrorString(e));
    if (e2 != cudaSuccess) return set_err(PLB_ERR_CUDA, "plb_site_genotypes_host: %s", cudaGetErrorString(e2));
    return PLB_OK;
}

// ---- S1 -----------------------------------------------------------------------------------------

namespace plb {
// explicit (read, segment) alignments; the packed path for reads >= 9 bp over ACGTN segments
__global__ void __launch_bounds__(128) k_align_batch(int n, const int64_t* __restrict__ hap_off,
                                                     const uint8_t* __restrict__ hap, const uint8_t* __restrict__ go,
                                                     const int64_t* __restrict__ read_off,
                                                     const uint8_t* __restrict__ rs, const uint8_t* __restrict__ rq,
                                                     u32* __restrict__ prof_ws, HapRec* __restrict__ rec_ws,
                                                     const int64_t* __restrict__ prof_off,
                                                     const int64_t* __restrict__ rec_off, int ext, int nuc,
                                                     int32_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int L = (int)(read_off[i + 1] - read_off[i]);
    const uint8_t* h = hap + hap_off[i];
    const uint8_t* g = go + hap_off[i];
    const uint8_t* r = rs + read_off[i];
    const uint8_t* q = rq + read_off[i];
    const int seg = L + 15;
    bool fast = L >= kMinFastLen && L <= kMaxFastLen;
    for (int x = 0; x < seg && fast; ++x) fast = fast_code(h[x]) != 5;
    bool six = fast;
    for (int x = 0; x < seg && six; ++x) six = fast_code(h[x]) != 4;
    bool five = six;   // every gap-open >= ext: one three-input min per cell pair
    for (int x = 0; x < seg && five; ++x) five = (int)g[x] >= ext;
    int v;
    if (fast) {
        u32* prof = prof_ws + prof_off[i];
        HapRec* rec = rec_ws + rec_off[i];
        const int n_rows = dp_steps(L) + 4;
        const int K = 2 * ext + nuc;
        for (int y = 0; y < n_rows; ++y) {
            u32 pv = 0u;
            if (y < L) pv = six ? make_profile6(fast_code(r[y]), q[y], K) : make_profile(fast_code(r[y]), q[y]);
            prof[y] = pv;
        }
        for (int x = 0; x < seg + kRecPad; ++x) {
            const int ca = x < seg ? fast_code(h[x]) : 4, cb = x + 4 < seg ? fast_code(h[x + 4]) : 4;
            const u32 oa = x < seg ? g[x] : 0u, ob = x + 4 < seg ? g[x + 4] : 0u;
            if (six) {
                rec[x].gow = pack_s16x2((int)oa - ext, (int)ob - ext);
                rec[x].sel = make_sel6(ca < 4 ? ca : 0, cb < 4 ? cb : 0);
            } else {
                rec[x].gow = oa | (ob << 16);
                rec[x].sel = make_sel(ca, cb);
            }
        }
        v = five ? band_dp_fast5(prof, rec, L, ext, nuc)
                 : six ? band_dp_fast6(prof, rec, L, ext, nuc) : band_dp_fast(prof, rec, L, ext, nuc);
    } else {
        v = band_dp_general(h, g, r, q, L, ext, nuc);
    }
    out[i] = v;
}

__global__ void k_gap_open_only(int n_haps, const int64_t* __restrict__ off, const uint8_t* __restrict__ seq,
                                uint8_t* __restrict__ out) {
    const int h = blockIdx.x;
    const int len = (int)(off[h + 1] - off[h]);
    const uint8_t* hap = seq + off[h];
    uint8_t* go = out + off[h] + h;
    for (int i = threadIdx.x; i <= len; i += blockDim.x) go[i] = i < len ? gap_open_at(hap, len, i) : 0;
}
}  // namespace plb

extern "C" int plb_align_batch_host(PlbContext* c, int32_t n, const int64_t* hap_seg_off, const uint8_t* hap_seg,
                                    const uint8_t* gap_open, const int64_t* read_off, const uint8_t* read_seq,
                                    const uint8_t* read_qual, int ext, int nuc, int32_t* scores_out) {
    if (!c || n < 0 || (n > 0 && (!hap_seg_off || !hap_seg || !gap_open || !read_off || !read_seq || !read_qual || !scores_out)))
        return set_err(PLB_ERR_ARG, "NULL / bad argument");
    if (n == 0) return PLB_OK;
    CU(cudaSetDevice(c->device));
    std::vector<int64_t> poff((size_t)n + 1), roff((size_t)n + 1);
    int64_t pw = 0, rw = 0;
    for (int i = 0; i < n; ++i) {
        const int64_t L = read_off[i + 1] - read_off[i];
        if (L < 1 || L > 32767) return set_err(PLB_ERR_SHAPE, "alignment %d: read length %lld out of range", i, (long long)L);
        if (hap_seg_off[i + 1] - hap_seg_off[i] < L + 15)
            return set_err(PLB_ERR_SHAPE, "alignment %d: segment shorter than read+15 (align.c:88)", i);
        poff[i] = pw;
        roff[i] = rw;
        pw += dp_steps((int)L) + 8;
        rw += L + 15 + kRecPad;
    }
    poff[n] = pw;
    roff[n] = rw;
    const int64_t hb = hap_seg_off[n], rb = read_off[n];
    Layout L;
    const size_t o_ho = L.take((size_t)(n + 1) * 8), o_h = L.take((size_t)hb + 64), o_g = L.take((size_t)hb + 64),
                 o_ro = L.take((size_t)(n + 1) * 8), o_rs = L.take((size_t)rb + 64), o_rq = L.take((size_t)rb + 64),
                 o_po = L.take((size_t)(n + 1) * 8), o_rco = L.take((size_t)(n + 1) * 8), o_pw = L.take((size_t)pw * 4),
                 o_rw = L.take((size_t)rw * sizeof(HapRec)), o_out = L.take((size_t)n * 4);
    Block B;
    int rc = block_get(c, L.off + 256, &B);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    cudaError_t e = cudaSuccess;
    auto up = [&](size_t off, const void* src, size_t bytes) {
        if (e == cudaSuccess && bytes) e = cudaMemcpyAsync((uint8_t*)B.p + off, src, bytes, cudaMemcpyHostToDevice, st);
    };
    up(o_ho, hap_seg_off, (size_t)(n + 1) * 8);
    up(o_h, hap_seg, (size_t)hb);
    up(o_g, gap_open, (size_t)hb);
    up(o_ro, read_off, (size_t)(n + 1) * 8);
    up(o_rs, read_seq, (size_t)rb);
    up(o_rq, read_qual, (size_t)rb);
    up(o_po, poff.data(), (size_t)(n + 1) * 8);
    up(o_rco, roff.data(), (size_t)(n + 1) * 8);
    if (e == cudaSuccess) {
        k_align_batch<<<(n + 127) / 128, 128, 0, st>>>(n, at<int64_t>(B, o_ho), at<uint8_t>(B, o_h), at<uint8_t>(B, o_g),
                                                        at<int64_t>(B, o_ro), at<uint8_t>(B, o_rs), at<uint8_t>(B, o_rq),
                                                        at<u32>(B, o_pw), at<HapRec>(B, o_rw), at<int64_t>(B, o_po),
                                                        at<int64_t>(B, o_rco), ext, nuc, at<int32_t>(B, o_out));
        e = cudaGetLastError();
        c->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(scores_out, (uint8_t*)B.p + o_out, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    block_put(c, B);
    if (e != cudaSuccess) return set_err(PLB_ERR_CUDA, "plb_align_batch_host: %s", cudaGetErrorString(e));
    if (e2 != cudaSuccess) return set_err(PLB_ERR_CUDA, "plb_align_batch_host: %s", cudaGetErrorString(e2));
    return PLB_OK;
}

namespace plb {
// scope row a2: explicit alignments with traceback rows (align.c:523-577) ...
__global__ void __launch_bounds__(64) k_align_traceback(int n, const int64_t* __restrict__ hap_off,
                                                        const uint8_t* __restrict__ hap, const uint8_t* __restrict__ go,
                                                        const int64_t* __restrict__ read_off,
                                                        const uint8_t* __restrict__ rs, const uint8_t* __restrict__ rq,
                                                        uint8_t* __restrict__ ptr_ws, char* __restrict__ aln1,
                                                        char* __restrict__ aln2, int ext, int nuc,
                                                        int32_t* __restrict__ score, int32_t* __restrict__ firstpos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t ro = read_off[i];
    const int L = (int)(read_off[i + 1] - ro);
    // per alignment: 16 back-pointer bytes per read row; alignment rows of 2L+16 bytes
    int fp = 0;
    score[i] = band_dp_traceback(hap + hap_off[i], go + hap_off[i], rs + ro, rq + ro, L, ext, nuc, ptr_ws + 16 * ro,
                                 aln1 + 2 * ro + 16 * (int64_t)i, aln2 + 2 * ro + 16 * (int64_t)i, &fp);
    firstpos[i] = fp;
}

// ... and with the flank score of calculateFlankScore (align.c:593-644) from one forward pass
__global__ void __launch_bounds__(64) k_align_flank(int n, const int64_t* __restrict__ hap_off,
                                                    const uint8_t* __restrict__ hap, const uint8_t* __restrict__ go,
                                                    const int32_t* __restrict__ seg_start,
                                                    const int32_t* __restrict__ hap_flank,
                                                    const int64_t* __restrict__ read_off, const uint8_t* __restrict__ rs,
                                                    const uint8_t* __restrict__ rq, int ext, int nuc,
                                                    int32_t* __restrict__ score, int32_t* __restrict__ flank) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t ro = read_off[i], ho = hap_off[i];
    const int L = (int)(read_off[i + 1] - ro), hap_len = (int)(hap_off[i + 1] - ho), st = seg_start[i];
    int fl = 0;
    score[i] = band_dp_flank(hap + ho + st, go + ho + st, rs + ro, rq + ro, L, ext, nuc, st, hap_len, hap_flank[i], &fl);
    flank[i] = fl;
}
}  // namespace plb

extern "C" int plb_align_traceback_host(PlbContext* c, int32_t n, const int64_t* hap_seg_off, const uint8_t* hap_seg,
                                        const uint8_t* gap_open, const int64_t* read_off, const uint8_t* read_seq,
                                        const uint8_t* read_qual, int ext, int nuc, int32_t* scores_out, char* aln1_out,
                                        char* aln2_out, int32_t* firstpos_out) {
    if (!c || n < 0 || (n > 0 && (!hap_seg_off || !hap_seg || !gap_open || !read_off || !read_seq || !read_qual ||
                                  !scores_out || !aln1_out || !aln2_out || !firstpos_out)))
        return set_err(PLB_ERR_ARG, "NULL / bad argument");
    if (n == 0) return PLB_OK;
    CU(cudaSetDevice(c->device));
    for (int i = 0; i < n; ++i) {
        const int64_t L = read_off[i + 1] - read_off[i];
        if (L < 1 || L > 32767) return set_err(PLB_ERR_SHAPE, "alignment %d: read length %lld out of range", i, (long long)L);
        if (hap_seg_off[i + 1] - hap_seg_off[i] < L + 15)
            return set_err(PLB_ERR_SHAPE, "alignment %d: segment shorter than read+15 (align.c:88)", i);
    }
    const int64_t hb = hap_seg_off[n], rb = read_off[n];
    const size_t aln_bytes = (size_t)(2 * rb + 16 * (int64_t)n);
    Layout L;
    const size_t o_ho = L.take((size_t)(n + 1) * 8), o_h = L.take((size_t)hb + 64), o_g = L.take((size_t)hb + 64),
                 o_ro = L.take((size_t)(n + 1) * 8), o_rs = L.take((size_t)rb + 64), o_rq = L.take((size_t)rb + 64),
                 o_ptr = L.take((size_t)rb * 16 + 64), o_a1 = L.take(aln_bytes), o_a2 = L.take(aln_bytes),
                 o_sc = L.take((size_t)n * 4), o_fp = L.take((size_t)n * 4);
    Block B;
    int rc = block_get(c, L.off + 256, &B);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    cudaError_t e = cudaSuccess;
    auto up = [&](size_t off, const void* src, size_t bytes) {
        if (e == cudaSuccess && bytes) e = cudaMemcpyAsync((uint8_t*)B.p + off, src, bytes, cudaMemcpyHostToDevice, st);
    };
    up(o_ho, hap_seg_off, (size_t)(n + 1) * 8);
    up(o_h, hap_seg, (size_t)hb);
    up(o_g, gap_open, (size_t)hb);
    up(o_ro, read_off, (size_t)(n + 1) * 8);
    up(o_rs, read_seq, (size_t)rb);
    up(o_rq, read_qual, (size_t)rb);
    if (e == cudaSuccess) {
        k_align_traceback<<<(n + 63) / 64, 64, 0, st>>>(n, at<int64_t>(B, o_ho), at<uint8_t>(B, o_h), at<uint8_t>(B, o_g),
                                                        at<int64_t>(B, o_ro), at<uint8_t>(B, o_rs), at<uint8_t>(B, o_rq),
                                                        at<uint8_t>(B, o_ptr), at<char>(B, o_a1), at<char>(B, o_a2), ext, nuc,
                                                        at<int32_t>(B, o_sc), at<int32_t>(B, o_fp));
        e = cudaGetLastError();
        c->launches++;
    }
    auto down = [&](void* dst, size_t off, size_t bytes) {
        if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(dst, (uint8_t*)B.p + off, bytes, cudaMemcpyDeviceToHost, st);
    };
    down(scores_out, o_sc, (size_t)n * 4);
    down(firstpos_out, o_fp, (size_t)n * 4);
    down(aln1_out, o_a1, aln_bytes);
    down(aln2_out, o_a2, aln_bytes);
    cudaError_t e2 = cudaStreamSynchronize(st);
    block_put(c, B);
    if (e != cudaSuccess) return set_err(PLB_ERR_CUDA, "plb_align_traceback_host: %s", cudaGetErrorString(e));
    if (e2 != cudaSuccess) return set_err(PLB_ERR_CUDA, "plb_align_traceback_host: %s", cudaGetErrorString(e2));
    return PLB_OK;
}

extern "C" int plb_align_flank_batch_host(PlbContext* c, int32_t n, const int64_t* hap_off, const uint8_t* hap_seq,
                                          const uint8_t* gap_open, const int32_t* seg_start, const int32_t* hap_flank,
                                          const int64_t* read_off, const uint8_t* read_seq, const uint8_t* read_qual,
                                          int ext, int nuc, int32_t* scores_out, int32_t* flank_out) {
    if (!c || n < 0 || (n > 0 && (!hap_off || !hap_seq || !gap_open || !seg_start || !hap_flank || !read_off ||
                                  !read_seq || !read_qual || !scores_out || !flank_out)))
        return set_err(PLB_ERR_ARG, "NULL / bad argument");
    if (n == 0) return PLB_OK;
    CU(cudaSetDevice(c->device));
    for (int i = 0; i < n; ++i) {
        const int64_t L = read_off[i + 1] - read_off[i];
        if (L < 1 || L > 32767) return set_err(PLB_ERR_SHAPE, "alignment %d: read length %lld out of range", i, (long long)L);
        if (seg_start[i] < 0 || hap_off[i + 1] - hap_off[i] < seg_start[i] + L + 15)
            return set_err(PLB_ERR_SHAPE, "alignment %d: haplotype shorter than start+read+15 (align.c:88)", i);
    }
    const int64_t hb = hap_off[n], rb = read_off[n];
    Layout L;
    const size_t o_ho = L.take((size_t)(n + 1) * 8), o_h = L.take((size_t)hb + 64), o_g = L.take((size_t)hb + 64),
                 o_ss = L.take((size_t)n * 4), o_hf = L.take((size_t)n * 4), o_ro = L.take((size_t)(n + 1) * 8),
                 o_rs = L.take((size_t)rb + 64), o_rq = L.take((size_t)rb + 64), o_sc = L.take((size_t)n * 4),
                 o_fl = L.take((size_t)n * 4);
    Block B;
    int rc = block_get(c, L.off + 256, &B);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    cudaError_t e = cudaSuccess;
    auto up = [&](size_t off, const void* src, size_t bytes) {
        if (e == cudaSuccess && bytes) e = cudaMemcpyAsync((uint8_t*)B.p + off, src, bytes, cudaMemcpyHostToDevice, st);
    };
    up(o_ho, hap_off, (size_t)(n + 1) * 8);
    up(o_h, hap_seq, (size_t)hb);
    up(o_g, gap_open, (size_t)hb);
    up(o_ss, seg_start, (size_t)n * 4);
    up(o_hf, hap_flank, (size_t)n * 4);
    up(o_ro, read_off, (size_t)(n + 1) * 8);
    up(o_rs, read_seq, (size_t)rb);
    up(o_rq, read_qual, (size_t)rb);
    if (e == cudaSuccess) {
        k_align_flank<<<(n + 63) / 64, 64, 0, st>>>(n, at<int64_t>(B, o_ho), at<uint8_t>(B, o_h), at<uint8_t>(B, o_g),
                                                    at<int32_t>(B, o_ss), at<int32_t>(B, o_hf), at<int64_t>(B, o_ro),
                                                    at<uint8_t>(B, o_rs), at<uint8_t>(B, o_rq), ext, nuc,
                                                    at<int32_t>(B, o_sc), at<int32_t>(B, o_fl));
        e = cudaGetLastError();
        c->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(scores_out, (uint8_t*)B.p + o_sc, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(flank_out, (uint8_t*)B.p + o_fl, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    block_put(c, B);
    if (e != cudaSuccess) return set_err(PLB_ERR_CUDA, "plb_align_flank_batch_host: %s", cudaGetErrorString(e));
    if (e2 != cudaSuccess) return set_err(PLB_ERR_CUDA, "plb_align_flank_batch_host: %s", cudaGetErrorString(e2));
    return PLB_OK;
}

extern "C" int plb_fast_align(PlbContext* c, const char* seq1, const char* seq2, const char* qual2, int len1, int len2,
                              int gapextend, int nucprior, const char* localgapopen, char* aln1, char* aln2,
                              int* firstpos) {
    if (!c || !seq1 || !seq2 || !qual2 || !localgapopen) return set_err(PLB_ERR_ARG, "NULL argument");
    if ((aln1 == nullptr) != (aln2 == nullptr)) return set_err(PLB_ERR_ARG, "aln1 and aln2 must both be given or both be NULL");
    if (len1 != len2 + 15) return set_err(PLB_ERR_SHAPE, "len1 must be len2 + 15 (align.c:88)");
    int64_t ho[2] = {0, len1}, ro[2] = {0, len2};
    int32_t score = 0;
    int rc;
    if (aln1) {  // traceback requested exactly as the reference decides it (align.c:96)
        int32_t fp = 0;
        rc = plb_align_traceback_host(c, 1, ho, (const uint8_t*)seq1, (const uint8_t*)localgapopen, ro, (const uint8_t*)seq2,
                                      (const uint8_t*)qual2, gapextend, nucprior, &score, aln1, aln2, &fp);
        if (rc == PLB_OK && firstpos) *firstpos = fp;
    } else {
        rc = plb_align_batch_host(c, 1, ho, (const uint8_t*)seq1, (const uint8_t*)localgapopen, ro, (const uint8_t*)seq2,
                                  (const uint8_t*)qual2, gapextend, nucprior, &score);
    }
    return rc ? rc : score;
}

extern "C" int plb_gap_open_host(PlbContext* c, int32_t n_haps, const int64_t* off, const uint8_t* seq, uint8_t* out) {
    if (!c || n_haps < 0 || (n_haps > 0 && (!off || !seq || !out))) return set_err(PLB_ERR_ARG, "NULL / bad argument");
    if (n_haps == 0) return PLB_OK;
    CU(cudaSetDevice(c->device));
    const int64_t nb = off[n_haps];
    Layout L;
    const size_t o_off = L.take((size_t)(n_haps + 1) * 8), o_seq = L.take((size_t)nb + 64), o_out = L.take((size_t)nb + n_haps + 64);
    Block B;
    int rc = block_get(c, L.off + 256, &B);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    cudaError_t e = cudaMemcpyAsync(at<uint8_t>(B, o_off), off, (size_t)(n_haps + 1) * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && nb) e = cudaMemcpyAsync(at<uint8_t>(B, o_seq), seq, (size_t)nb, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        k_gap_open_only<<<n_haps, 128, 0, st>>>(n_haps, at<int64_t>(B, o_off), at<uint8_t>(B, o_seq), at<uint8_t>(B, o_out));
        e = cudaGetLastError();
        c->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, at<uint8_t>(B, o_out), (size_t)nb + n_haps, cudaMemcpyDeviceToHost, st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    block_put(c, B);
    if (e != cudaSuccess) return set_err(PLB_ERR_CUDA, "plb_gap_open_host: %s", cudaGetErrorString(e));
    if (e2 != cudaSuccess) return set_err(PLB_ERR_CUDA, "plb_gap_open_host: %s", cudaGetErrorString(e2));
    return PLB_OK;
}

// ---- N1: haplotype construction + selection loop ------------------------------------------------
#include "plb_select.cuh"

// ---- N3: read staging (host) ----------------------------------------------------------------------
#include "plb_stage.cuh"

// ---- measurement support: synthetic windows generated on the device (BASELINE config 5) -----------------
#include "plb_synth.cuh"
