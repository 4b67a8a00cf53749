// plb_select.cuh — SURVEY §8f row N1: haplotype construction and the haplotype selection loop.
// Included at the end of plb_api.cu (uses its static helpers).
//
// Reference: src/cython/variantFilter.pyx:377-506 getFilteredHaplotypes, :237-283 computeBestScoreForGenotype,
// src/cython/platypusutils.pyx:735-802 isHaplotypeValid, src/cython/chaplotype.pyx:127-172, 397-449
// (Haplotype.__init__ / getMutatedSequence), src/cython/variant.pyx:109-145, 282-353 (Variant fields and order).
//
// The reference scores one trial haplotype at a time against one window's reads, in nVar data-dependent rounds per
// window.  Here a round is ONE batch over every window that still has a variant to add: the trial haplotypes are
// built on the GPU from the window's reference segment and variant list (k_build_haps), every sampled read is scored
// against them with the S2 kernels (k_prep / k_anchor / k_general / k_dp; reads passed as broken mates = a bare
// alignReadToHaplotype = alignSingleRead), and k_trial_score reduces sum_r log(0.5 (e^LL_ref + e^LL_trial)) per
// (trial, individual).  Reads, slots and the reference-haplotype log-likelihoods stay on the GPU for all rounds; per
// round only the trial masks go up and one double per trial comes back, and the host replays the reference's heap
// bookkeeping on those scores.

namespace plb {

struct SelVars {   // variants of the windows of a selection / construction call, on host or device
    const int32_t* var_off;        // [W+1]
    const int32_t* var_pos;
    const int32_t* var_nrem;
    const int64_t* var_added_off;  // [n_vars+1]
    const uint8_t* var_added;
};

// Haplotype.haplotypeSequence as a list of pieces (chaplotype.pyx:163-172, 397-449).  emit(kind, a, n): kind 0 =
// n reference bases from genomic position a, kind 1 = the n added bases of variant a (batch-wide index).
// ref_start / ref_len: genomic position and length of the window's reference segment
// (= refFile.getSequence(startPos - endBufferSize, endPos + endBufferSize), clamped to the contig).
template <typename Emit>
__host__ __device__ __forceinline__ void walk_haplotype(int win_start, int win_end, int ref_start, int ref_len,
                                                        uint64_t mask, int v0, const int32_t* pos, const int32_t* nrem,
                                                        const int64_t* added_off, Emit&& emit) {
    const int ref_end = ref_start + ref_len;
    auto ref = [&](int a, int b) {   // refFile.getSequence(a, b) inside the segment
        if (a < ref_start) a = ref_start;
        if (b > ref_end) b = ref_end;
        if (b > a) emit(0, a, b - a);
    };
    if (mask == 0) {   // no variants: the reference sequence itself
        ref(ref_start, ref_end);
        return;
    }
    ref(ref_start, win_start);   // leftBuffer
    int cur = win_start;
    for (uint64_t m = mask; m; m &= m - 1) {
#ifdef __CUDA_ARCH__
        const int v = v0 + __ffsll((long long)m) - 1;
#else
        const int v = v0 + __builtin_ctzll(m);
#endif
        const int p = pos[v], nr = nrem[v], na = (int)(added_off[v + 1] - added_off[v]);
        if (p > cur) {   // reference up to the variant (also the "sequence before the first variant" case)
            ref(cur, p);
            cur = p;
        }
        if (na == nr) {   // SNP / MNP
            emit(1, v, na);
            cur += nr;
        } else {
            if (na == 0 || nr == 0) {   // insertion / deletion: the anchor base, unless already passed
                if (p == cur) {
                    ref(p, p + 1);
                    cur += 1;
                }
            }
            cur += nr;
            emit(1, v, na);
        }
    }
    if (cur < win_end) ref(cur, win_end);
    ref(win_end, ref_end);   // rightBuffer
}

constexpr int kBuildWarps = 4;
// walk_haplotype emits up to three pieces per variant (reference stretch, anchor base, added bases) plus the two flanks
// and the window's tail: 3 * 64 + 3 for the 64 variants a mask can name
constexpr int kBuildMaxPieces = 3 * 64 + 3;

// One warp per haplotype: lane 0 lists the pieces, all lanes copy them.
__global__ void __launch_bounds__(32 * kBuildWarps) k_build_haps(int n_haps, const int32_t* __restrict__ hap_win,
                                                                const uint64_t* __restrict__ hap_mask,
                                                                const int64_t* __restrict__ ref_off,
                                                                const uint8_t* __restrict__ ref_seq,
                                                                const int32_t* __restrict__ win_start,
                                                                const int32_t* __restrict__ win_end,
                                                                const int32_t* __restrict__ hap_start, SelVars sv,
                                                                const int64_t* __restrict__ out_off,
                                                                uint8_t* __restrict__ out) {
    __shared__ const uint8_t* s_src[kBuildWarps][kBuildMaxPieces];
    __shared__ int s_len[kBuildWarps][kBuildMaxPieces];
    __shared__ int s_n[kBuildWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int h = blockIdx.x * kBuildWarps + warp; h < n_haps; h += gridDim.x * kBuildWarps) {
        const int w = hap_win[h];
        const int ws = win_start[w], we = win_end[w];
        const int left = min(ws - hap_start[w], ws);
        const int ref_start = ws - left;
        const uint8_t* rs = ref_seq + ref_off[w];
        if (lane == 0) {
            int n = 0;
            walk_haplotype(ws, we, ref_start, (int)(ref_off[w + 1] - ref_off[w]), hap_mask[h], sv.var_off[w], sv.var_pos,
                           sv.var_nrem, sv.var_added_off, [&](int kind, int a, int len) {
                               s_src[warp][n] = kind ? sv.var_added + sv.var_added_off[a] : rs + (a - ref_start);
                               s_len[warp][n] = len;
                               ++n;
                           });
            s_n[warp] = n;
        }
        __syncwarp();
        uint8_t* dst = out + out_off[h];
        const int n = s_n[warp];
        for (int k = 0; k < n; ++k) {
            const uint8_t* src = s_src[warp][k];
            const int len = s_len[warp][k];
            for (int i = lane; i < len; i += 32) dst[i] = src[i];
            dst += len;
        }
        __syncwarp();
    }
}

// computeBestScoreForGenotype (variantFilter.pyx:237-283) for every trial haplotype of a round: one warp per
// haplotype.  ll = the round's per-read log-likelihoods (PlbLoglikOut layout), ll_ref[slot] = the same reads
// against the window's reference haplotype.  Lanes evaluate log(0.5 (e^s1 + e^s2)) for 32 reads at a time; the terms
// are then added in read order, as the reference adds them.
__global__ void __launch_bounds__(128) k_trial_score(DevBatch b, const double* __restrict__ ll,
                                                     const double* __restrict__ ll_ref, int n_haps,
                                                     double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int h = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (h >= n_haps) return;
    const int w = b.hap_win[h];
    const int hl = h - b.win_hap_off[w];
    const int nInd = b.n_individuals;
    double best = -1e20;
    for (int i = 0; i < nInd; ++i) {
        const int64_t wi = (int64_t)w * nInd + i;
        const int64_t s0 = b.wi_slot_off[wi];
        const int T = (int)(b.wi_slot_off[wi + 1] - s0);
        if (T == 0) continue;   // readBegin == readEnd
        const double* l2 = ll + b.ll_off[wi] + (int64_t)hl * T;
        const double* l1 = ll_ref + s0;
        double tot = 0.0;
        for (int t0 = 0; t0 < T; t0 += 32) {
            const int t = t0 + lane;
            double term = 0.0;
            if (t < T) term = log(0.5 * (exp(l1[t]) + exp(l2[t])));
            const int n = min(32, T - t0);
            for (int k = 0; k < n; ++k) tot += __shfl_sync(0xffffffffu, term, k);
        }
        if (tot > best) best = tot;
    }
    if (lane == 0) out[h] = best;
}

// computeBestScoreForHaplotype (variantFilter.pyx:212-234) for every haplotype of a batch whose slots are the good reads:
// per individual the per-read log-likelihoods are added in read order, best individual (one without reads sums to 0.0).
__global__ void __launch_bounds__(128) k_hap_score(DevBatch b, const double* __restrict__ ll, int n_haps,
                                                   double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int h = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (h >= n_haps) return;
    const int w = b.hap_win[h];
    const int hl = h - b.win_hap_off[w];
    const int nInd = b.n_individuals;
    double best = -1e20;
    for (int i = 0; i < nInd; ++i) {
        const int64_t wi = (int64_t)w * nInd + i;
        const int T = (int)(b.wi_slot_off[wi + 1] - b.wi_slot_off[wi]);
        const double* l = ll + b.ll_off[wi] + (int64_t)hl * T;
        double tot = 0.0;
        for (int t0 = 0; t0 < T; t0 += 32) {
            const int t = t0 + lane;
            const double term = t < T ? l[t] : 0.0;
            const int n = min(32, T - t0);
            for (int k = 0; k < n; ++k) tot += __shfl_sync(0xffffffffu, term, k);
        }
        if (tot > best) best = tot;
    }
    if (lane == 0) out[h] = best;
}

// computeBestScoreForGenotype (variantFilter.pyx:237-283) for arbitrary pairs of haplotypes of one window: one warp per
// pair, the batch's slots are the sampled good reads.  Same arithmetic and order of additions as k_trial_score.
__global__ void __launch_bounds__(128) k_pair_score(DevBatch b, const double* __restrict__ ll, int n_pairs,
                                                    const int32_t* __restrict__ hap1, const int32_t* __restrict__ hap2,
                                                    double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n_pairs) return;
    const int h1 = hap1[p], h2 = hap2[p];
    const int w = b.hap_win[h1];
    const int l1h = h1 - b.win_hap_off[w], l2h = h2 - b.win_hap_off[w];
    const int nInd = b.n_individuals;
    double best = -1e20;
    for (int i = 0; i < nInd; ++i) {
        const int64_t wi = (int64_t)w * nInd + i;
        const int T = (int)(b.wi_slot_off[wi + 1] - b.wi_slot_off[wi]);
        if (T == 0) continue;   // readBegin == readEnd
        const double* l1 = ll + b.ll_off[wi] + (int64_t)l1h * T;
        const double* l2 = ll + b.ll_off[wi] + (int64_t)l2h * T;
        double tot = 0.0;
        for (int t0 = 0; t0 < T; t0 += 32) {
            const int t = t0 + lane;
            double term = 0.0;
            if (t < T) term = log(0.5 * (exp(l1[t]) + exp(l2[t])));
            const int n = min(32, T - t0);
            for (int k = 0; k < n; ++k) tot += __shfl_sync(0xffffffffu, term, k);
        }
        if (tot > best) best = tot;
    }
    if (lane == 0) out[p] = best;
}

}  // namespace plb

namespace {

// ---- the reference's bookkeeping on (score, variants) entries -------------------------------------------------

struct SelEntry {
    double score;
    uint64_t mask;   // the variant tuple: always in the window's variant order (see sel_lt)
};

struct SelKeys {   // per window: the Variant order key (variant.pyx:304-315) of each of its variants
    const int32_t* pos;
    const int32_t* type;
    const int32_t* nrem;
    bool var_lt(int a, int b) const {
        if (pos[a] != pos[b]) return pos[a] < pos[b];
        if (type[a] != type[b]) return type[a] < type[b];
        return nrem[a] < nrem[b];
    }
};

// (score, variants) < (score, variants) as CPython compares tuples: the first pair of unequal items decides.  A
// variant tuple in the heap is a valid haplotype, so no two of its variants share an order key and the tuple is in
// ascending variant index; variants of one window are distinct, so "equal" is "same index".  Two different variants
// with the same key (two alleles of one SNP) are unequal and neither is less: the comparison is False both ways.
static bool sel_lt(const SelEntry& a, const SelEntry& b, const SelKeys& k) {
    if (a.score != b.score) return a.score < b.score;
    uint64_t ma = a.mask, mb = b.mask;
    while (ma && mb) {
        const int va = __builtin_ctzll(ma), vb = __builtin_ctzll(mb);
        if (va != vb) return k.var_lt(va, vb);
        ma &= ma - 1;
        mb &= mb - 1;
    }
    return ma == 0 && mb != 0;   // the shorter tuple is less
}

// heapq._siftdown / _siftup / heappush / heappushpop (CPython Lib/heapq.py = Modules/_heapqmodule.c)
static void heap_siftdown(std::vector<SelEntry>& h, int startpos, int pos, const SelKeys& k) {
    const SelEntry item = h[pos];
    while (pos > startpos) {
        const int parent = (pos - 1) >> 1;
        if (sel_lt(item, h[parent], k)) {
            h[pos] = h[parent];
            pos = parent;
            continue;
        }
        break;
    }
    h[pos] = item;
}

static void heap_siftup(std::vector<SelEntry>& h, int pos, const SelKeys& k) {
    const int end = (int)h.size(), start = pos;
    const SelEntry item = h[pos];
    int child = 2 * pos + 1;
    while (child < end) {
        const int right = child + 1;
        if (right < end && !sel_lt(h[child], h[right], k)) child = right;
        h[pos] = h[child];
        pos = child;
        child = 2 * pos + 1;
    }
    h[pos] = item;
    heap_siftdown(h, start, pos, k);
}

static void heap_push(std::vector<SelEntry>& h, const SelEntry& e, const SelKeys& k) {
    h.push_back(e);
    heap_siftdown(h, 0, (int)h.size() - 1, k);
}

static void heap_pushpop(std::vector<SelEntry>& h, SelEntry e, const SelKeys& k) {
    if (!h.empty() && sel_lt(h[0], e, k)) {
        std::swap(e, h[0]);
        heap_siftup(h, 0, k);
    }
}

// list.sort() for fewer than 64 items (CPython Objects/listobject.c: minrun = n, so the whole list is one
// count_run + binarysort); `reverse` = sort(reverse=True) (reverse, sort, reverse).  Written out because the order
// above is not a strict weak order (see sel_lt), so the result depends on the algorithm.
static void py_sort(std::vector<SelEntry>& a, const SelKeys& k, bool reverse) {
    const int n = (int)a.size();
    {   // distinct scores: the order is total and every correct sort returns the same list
        std::vector<SelEntry> b = a;
        std::sort(b.begin(), b.end(), [](const SelEntry& x, const SelEntry& y) { return x.score < y.score; });
        bool distinct = true;
        for (int i = 1; i < n && distinct; ++i) distinct = b[(size_t)i - 1].score != b[(size_t)i].score;
        if (distinct) {
            if (reverse) std::reverse(b.begin(), b.end());
            a.swap(b);
            return;
        }
    }
    if (reverse) std::reverse(a.begin(), a.end());
    if (n >= 2) {
        int run = 2;
        if (sel_lt(a[1], a[0], k)) {   // strictly descending run
            while (run < n && sel_lt(a[run], a[run - 1], k)) ++run;
            std::reverse(a.begin(), a.begin() + run);
        } else {
            while (run < n && !sel_lt(a[run], a[run - 1], k)) ++run;
        }
        for (int start = run; start < n; ++start) {   // binarysort
            int l = 0, r = start;
            const SelEntry pivot = a[start];
            do {
                const int p = l + ((r - l) >> 1);
                if (sel_lt(pivot, a[p], k))
                    r = p;
                else
                    l = p + 1;
            } while (l < r);
            for (int p = start; p > l; --p) a[p] = a[p - 1];
            a[l] = pivot;
        }
    }
    if (reverse) std::reverse(a.begin(), a.end());
}

// isHaplotypeValid (platypusutils.pyx:735-802) on a mask; variants in window order (sorted by key).
static bool sel_valid(uint64_t mask, const int32_t* pos, const int32_t* nrem, const int32_t* nadd) {
    int prev = -1;
    for (uint64_t m = mask; m; m &= m - 1) {
        const int v = __builtin_ctzll(m);
        if (prev >= 0) {
            const int a_max = std::max(pos[prev], pos[prev] + nrem[prev] - 1), b_min = pos[v];
            if (a_max > b_min) return false;
            if (a_max == b_min) {
                const bool a_snp = nadd[prev] == nrem[prev];
                if (!(a_snp && nadd[v] != nrem[v])) return false;
            }
        }
        prev = v;
    }
    return true;
}

static int var_type(int nrem, int nadd) {   // variant.pyx:136-144
    if (nrem == nadd) return nadd == 1 ? 0 : 1;
    if (nrem == 0) return 2;
    if (nadd == 0) return 3;
    return 4;
}

struct SelHost {   // validated per-variant host arrays of a call
    std::vector<int32_t> nadd, type;
};

static int check_variants(const PlbWindowBatch* rb, const PlbVariantSet* vs, SelHost& sh) {
    if (!rb || !vs || !vs->win_var_off) return set_err(PLB_ERR_ARG, "NULL argument");
    const int W = rb->n_windows;
    if (W < 0) return set_err(PLB_ERR_ARG, "bad n_windows");
    if (W == 0) return PLB_OK;
    if (!rb->win_hap_off || !rb->win_start || !rb->win_end || !rb->hap_start || !rb->hap_seq_off || !rb->hap_seq)
        return set_err(PLB_ERR_ARG, "ref_batch: NULL window / haplotype array");
    const int nv = vs->win_var_off[W];
    if (nv > 0 && (!vs->var_pos || !vs->var_n_removed || !vs->var_added_off || (vs->var_added_off[nv] > 0 && !vs->var_added)))
        return set_err(PLB_ERR_ARG, "PlbVariantSet: NULL array");
    sh.nadd.resize((size_t)nv);
    sh.type.resize((size_t)nv);
    for (int v = 0; v < nv; ++v) {
        const int64_t na = vs->var_added_off[v + 1] - vs->var_added_off[v];
        if (na < 0 || na > PLB_MAX_HAP_LEN || vs->var_n_removed[v] < 0)
            return set_err(PLB_ERR_ARG, "variant %d: bad added / removed length", v);
        sh.nadd[(size_t)v] = (int)na;
        sh.type[(size_t)v] = var_type(vs->var_n_removed[v], (int)na);
    }
    for (int w = 0; w < W; ++w) {
        if (rb->win_hap_off[w + 1] - rb->win_hap_off[w] != 1 || rb->win_hap_off[w] != w)
            return set_err(PLB_ERR_ARG, "ref_batch must hold exactly one (reference) haplotype per window (window %d)", w);
        const int v0 = vs->win_var_off[w], v1 = vs->win_var_off[w + 1];
        if (v1 < v0) return set_err(PLB_ERR_ARG, "win_var_off not monotone at window %d", w);
        if (v1 - v0 > 64) return set_err(PLB_ERR_SHAPE, "window %d has %d variants (max 64)", w, v1 - v0);
        const int ws = rb->win_start[w], we = rb->win_end[w];
        const int left = std::min(ws - rb->hap_start[w], ws);
        const int64_t ref_len = rb->hap_seq_off[w + 1] - rb->hap_seq_off[w];
        if (ws < 0 || we <= ws || left < 0 || ref_len < (int64_t)left + (we - ws))
            return set_err(PLB_ERR_ARG, "window %d: interval / reference segment inconsistent", w);
        for (int v = v0; v < v1; ++v) {
            const int p = vs->var_pos[v];
            if (p < ws || p > we || p >= ws - left + ref_len)
                return set_err(PLB_ERR_ARG, "variant %d of window %d lies outside [win_start, win_end]", v - v0, w);
            if (v > v0) {
                const int q = v - 1;
                const bool lt = vs->var_pos[q] != p ? vs->var_pos[q] < p
                                : sh.type[(size_t)q] != sh.type[(size_t)v] ? sh.type[(size_t)q] < sh.type[(size_t)v]
                                : vs->var_n_removed[q] <= vs->var_n_removed[v];
                if (!lt) return set_err(PLB_ERR_ARG, "variants of window %d are not sorted (variant.pyx:304-315)", w);
            }
            for (int q = v - 1; q >= v0 && vs->var_pos[q] == p; --q)   // duplicates: same position, lengths and added bases
                if (vs->var_n_removed[q] == vs->var_n_removed[v] && sh.nadd[(size_t)q] == sh.nadd[(size_t)v] &&
                    memcmp(vs->var_added + vs->var_added_off[q], vs->var_added + vs->var_added_off[v], (size_t)sh.nadd[(size_t)v]) == 0)
                    return set_err(PLB_ERR_ARG, "window %d lists a variant twice", w);
        }
    }
    return PLB_OK;
}

static int64_t hap_length(const PlbWindowBatch* rb, const PlbVariantSet* vs, int w, uint64_t mask) {
    const int ws = rb->win_start[w], we = rb->win_end[w];
    const int left = std::min(ws - rb->hap_start[w], ws);
    int64_t n = 0;
    walk_haplotype(ws, we, ws - left, (int)(rb->hap_seq_off[w + 1] - rb->hap_seq_off[w]), mask, vs->win_var_off[w],
                   vs->var_pos, vs->var_n_removed, vs->var_added_off, [&](int, int, int len) { n += len; });
    return n;
}

struct SelStats {
    double v[10];
};
static thread_local SelStats g_sel_stats{};

}  // namespace

extern "C" int plb_select_stats(PlbContext*, double* out, int n) {   // the statistics are per host thread
    if (!out || n < 0) return set_err(PLB_ERR_ARG, "NULL argument");
    for (int i = 0; i < n; ++i) out[i] = i < 10 ? g_sel_stats.v[i] : 0.0;
    return PLB_OK;
}

extern "C" int plb_build_haplotypes_host(PlbContext* c, const PlbWindowBatch* rb, const PlbVariantSet* vs, int32_t n_haps,
                                         const int32_t* hap_win, const uint64_t* hap_mask, int64_t* hap_seq_off,
                                         uint8_t* hap_seq, int64_t capacity) {
    // ctx may be NULL when only the offsets are wanted (hap_seq == NULL): the lengths are computed on the host
    if ((!c && hap_seq) || n_haps < 0 || (n_haps > 0 && (!hap_win || !hap_mask)) || !hap_seq_off)
        return set_err(PLB_ERR_ARG, "NULL / bad argument");
    if (rb && rb->seq_format != PLB_SEQ_ASCII)
        return set_err(PLB_ERR_UNSUPPORTED, "plb_build_haplotypes_host takes ASCII reference segments only");
    SelHost sh;
    int rc = check_variants(rb, vs, sh);
    if (rc) return rc;
    const int W = rb->n_windows;
    hap_seq_off[0] = 0;
    for (int h = 0; h < n_haps; ++h) {
        const int w = hap_win[h];
        if (w < 0 || w >= W) return set_err(PLB_ERR_ARG, "hap_win[%d] out of range", h);
        const int nv = vs->win_var_off[w + 1] - vs->win_var_off[w];
        if (nv < 64 && (hap_mask[h] >> nv)) return set_err(PLB_ERR_ARG, "hap_mask[%d] names a variant the window does not have", h);
        const int64_t len = hap_length(rb, vs, w, hap_mask[h]);
        if (len > PLB_MAX_HAP_LEN)
            return set_err(PLB_ERR_SHAPE, "haplotype %d has length %lld (max %d, chaplotype.pyx:180-183)", h, (long long)len, PLB_MAX_HAP_LEN);
        hap_seq_off[h + 1] = hap_seq_off[h] + len;
    }
    if (!hap_seq || n_haps == 0) return PLB_OK;
    const int64_t out_bytes = hap_seq_off[n_haps];
    if (out_bytes > capacity) return set_err(PLB_ERR_ARG, "hap_seq holds %lld bytes, %lld needed", (long long)capacity, (long long)out_bytes);
    CU(cudaSetDevice(c->device));
    const int nvar = vs->win_var_off[W];
    const int64_t ref_bytes = rb->hap_seq_off[W], add_bytes = nvar ? vs->var_added_off[nvar] : 0;
    Layout L;
    const size_t o_ws = L.take((size_t)W * 4), o_we = L.take((size_t)W * 4), o_hs = L.take((size_t)W * 4),
                 o_roff = L.take((size_t)(W + 1) * 8), o_ref = L.take((size_t)ref_bytes + 64),
                 o_voff = L.take((size_t)(W + 1) * 4), o_vpos = L.take((size_t)nvar * 4), o_vnr = L.take((size_t)nvar * 4),
                 o_vaoff = L.take((size_t)(nvar + 1) * 8), o_vadd = L.take((size_t)add_bytes + 64),
                 o_hw = L.take((size_t)n_haps * 4), o_hm = L.take((size_t)n_haps * 8), o_ooff = L.take((size_t)(n_haps + 1) * 8),
                 o_out = L.take((size_t)out_bytes + 64);
    Block B;
    if ((rc = block_get(c, L.off + 256, &B))) return rc;
    cudaStream_t st = c->stream;
    cudaError_t e = cudaSuccess;
    auto up = [&](size_t off, const void* src, size_t bytes) {
        if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(at<uint8_t>(B, off), src, bytes, cudaMemcpyHostToDevice, st);
    };
    const int64_t zero = 0;
    up(o_ws, rb->win_start, (size_t)W * 4);
    up(o_we, rb->win_end, (size_t)W * 4);
    up(o_hs, rb->hap_start, (size_t)W * 4);
    up(o_roff, rb->hap_seq_off, (size_t)(W + 1) * 8);
    up(o_ref, rb->hap_seq, (size_t)ref_bytes);
    up(o_voff, vs->win_var_off, (size_t)(W + 1) * 4);
    up(o_vpos, vs->var_pos, (size_t)nvar * 4);
    up(o_vnr, vs->var_n_removed, (size_t)nvar * 4);
    if (nvar)
        up(o_vaoff, vs->var_added_off, (size_t)(nvar + 1) * 8);
    else
        up(o_vaoff, &zero, 8);
    up(o_vadd, vs->var_added, (size_t)add_bytes);
    up(o_hw, hap_win, (size_t)n_haps * 4);
    up(o_hm, hap_mask, (size_t)n_haps * 8);
    up(o_ooff, hap_seq_off, (size_t)(n_haps + 1) * 8);
    if (e == cudaSuccess) {
        SelVars sv{at<int32_t>(B, o_voff), at<int32_t>(B, o_vpos), at<int32_t>(B, o_vnr), at<int64_t>(B, o_vaoff), at<uint8_t>(B, o_vadd)};
        const int grid = std::max(1, std::min((n_haps + kBuildWarps - 1) / kBuildWarps, c->n_sm * 16));
        k_build_haps<<<grid, 32 * kBuildWarps, 0, st>>>(n_haps, at<int32_t>(B, o_hw), at<uint64_t>(B, o_hm), at<int64_t>(B, o_roff),
                                                       at<uint8_t>(B, o_ref), at<int32_t>(B, o_ws), at<int32_t>(B, o_we),
                                                       at<int32_t>(B, o_hs), sv, at<int64_t>(B, o_ooff), at<uint8_t>(B, o_out));
        e = cudaGetLastError();
        c->launches++;
    }
    if (e == cudaSuccess && out_bytes) e = cudaMemcpyAsync(hap_seq, at<uint8_t>(B, o_out), (size_t)out_bytes, cudaMemcpyDeviceToHost, st);
    const cudaError_t e2 = cudaStreamSynchronize(st);
    block_put(c, B);
    if (e != cudaSuccess) return set_err(PLB_ERR_CUDA, "plb_build_haplotypes_host: %s", cudaGetErrorString(e));
    if (e2 != cudaSuccess) return set_err(PLB_ERR_CUDA, "plb_build_haplotypes_host: %s", cudaGetErrorString(e2));
    return PLB_OK;
}

// Length of a VALID haplotype without walking it: every variant changes the length by nAdded - nRemoved (anchor
// bases and reference stretches are copied one for one) unless the walk runs past the window end, where the right
// buffer would repeat bases; a valid set keeps `cur` <= pos + 1 before each variant, so cur never exceeds
// max(pos + nRemoved + 2).  Returns -1 when that bound crosses win_end (the caller walks instead).
static inline int64_t fast_hap_length(uint64_t mask, int64_t ref_len, int win_end, const int32_t* pos, const int32_t* nrem,
                                      const int32_t* nadd) {
    int64_t len = ref_len;
    for (uint64_t m = mask; m; m &= m - 1) {
        const int v = __builtin_ctzll(m);
        if (pos[v] + nrem[v] + 2 > win_end) return -1;
        len += nadd[v] - nrem[v];
    }
    return len;
}

// One round of one group of windows as the scorer sees it: Wr windows (a prefix of the group's processing order), nh
// trial haplotypes in window order.  The arrays live in pinned host memory when the scorer supplied it.
struct SelRound {
    int Wr, nh;
    const int32_t* hap_off;     // [Wr+1]
    const int64_t* hap_seq_off; // [nh+1] lengths computed on the host
    const uint64_t* mask;       // [nh]
    const int* n_trials;        // [Wr]
};

struct SelEntryState {
    std::vector<int> order;       // variants by decreasing nSupportingReads (stable), window-local indices
    std::vector<SelEntry> heap;
    int n_done = 0;
};

// Host-side inputs and state of the scoring rounds of one GROUP of windows: filter-branch windows in processing
// order (decreasing variant count, so that the windows of round r are a prefix).
struct SelPlan {
    int gid = 0, Wf = 0, nInd = 1, max_rounds = 0, max_trials = 0;
    std::vector<int> filt;   // processing order -> window of the caller's batch
    std::vector<int32_t> ws, we, hs, hoff, zero;
    std::vector<int64_t> hsoff, slot_off;
    std::vector<int32_t> slot;
    std::vector<uint8_t> ref;
    std::vector<int32_t> voff, pos, nrem, nadd, type, nsup;
    std::vector<int64_t> aoff;
    std::vector<uint8_t> add;
    std::vector<int64_t> win_cells;   // 16 * readLen summed over the window's sampled reads
    PlbWindowBatch batch{};           // one reference haplotype per window + the sampled reads as broken mates
    // rounds
    std::vector<SelEntryState> state;
    std::vector<uint64_t> trial_mask;
    std::vector<int> n_trials;
    std::vector<int64_t> win_base;    // scratch of the offset computation
    int r = 0, Wr = 0, nh = 0;
    bool active = false;
    bool in_prologue = false;   // the launch in flight scores the trial sets of rounds 0 .. prologue-1 together
    // per-round arrays handed to the scorer: pinned when the scorer provides memory (set in prepare), else own_*
    int32_t* r_hoff = nullptr;
    int64_t* r_hsoff = nullptr;
    uint64_t* mask_c = nullptr;
    double* scores = nullptr;
    std::vector<int32_t> own_hoff;
    std::vector<int64_t> own_hsoff;
    std::vector<uint64_t> own_mask;
    std::vector<double> own_scores;
};

// The reference's loop with the scoring of a round's trial haplotypes left to the scorer (the GPU in
// plb_select_haplotypes_host; the caller in plb_select_replay_host):
//   prepare(plan)        once per group before its rounds (may point plan.r_hoff / r_hsoff / mask_c / scores at pinned memory)
//   submit(plan, round)  start scoring the round's trial haplotypes into plan.scores (may return before they are there)
//   wait(plan)           block until plan.scores is complete
// Large batches are cut into two groups whose rounds alternate, so that the host's bookkeeping for one group runs
// while the scorer works on the other.
template <typename Alloc, typename Prepare, typename Submit, typename Wait>
static int select_core(const PlbWindowBatch* rb, const PlbVariantSet* vs, const PlbSelectOptions* so, const PlbOptions* opt,
                       PlbSelectOut* out, bool need_reads, Alloc&& alloc, Prepare&& prepare, Submit&& submit, Wait&& wait) {
    if (!so || !out || !out->n_sel || !out->sel_mask) return set_err(PLB_ERR_ARG, "NULL argument");
    SelHost sh;
    int rc = check_variants(rb, vs, sh);
    if (rc) return rc;
    const int W = rb->n_windows, nInd = rb->n_individuals;
    if (W == 0) return PLB_OK;
    if (nInd < 1 || (need_reads && (!rb->wi_slot_off || !rb->wi_n_good))) return set_err(PLB_ERR_ARG, "ref_batch: NULL slot arrays");
    if (need_reads && rb->n_slots > 0 && (!rb->slot_read || !rb->read_seq_off || !rb->read_seq || !rb->read_qual || !rb->read_pos ||
                                          !rb->read_end || !rb->read_mapq || !rb->read_qcfail))
        return set_err(PLB_ERR_ARG, "NULL read array in batch");
    const int orig_cap = so->original_max_haplotypes - 1, cap = so->max_haplotypes - 1;
    if (cap < 1 || orig_cap < 1 || so->coverage_sampling_level <= 0) return set_err(PLB_ERR_ARG, "bad PlbSelectOptions");
    if (orig_cap > 63)
        return set_err(PLB_ERR_SHAPE, "original_max_haplotypes - 1 = %d > 63 (list.sort beyond one run is not restated)", orig_cap);
    if (out->max_sel < 1) return set_err(PLB_ERR_ARG, "max_sel < 1");
    if (!vs->var_n_support && vs->win_var_off[W] > 0) return set_err(PLB_ERR_ARG, "var_n_support is NULL");
    memset(&g_sel_stats, 0, sizeof g_sel_stats);
    const double log2cap = std::log2((double)cap);
    const double nan_v = std::nan("");
    static const bool check_len = getenv("PLB_SELECT_CHECK") != nullptr;

    // ---- windows with few variants: every valid combination, in itertools.combinations order (:411-438)
    std::vector<int> filt;   // windows that take the scoring rounds
    for (int w = 0; w < W; ++w) {
        const int v0 = vs->win_var_off[w], n = vs->win_var_off[w + 1] - v0;
        if (out->n_scored) out->n_scored[w] = 0;
        const bool all = (double)n <= log2cap || (so->filter_vars_by_coverage && (double)so->max_variants <= log2cap);
        if (!all) {
            filt.push_back(w);
            continue;
        }
        int k_out = 0;
        uint64_t* masks = out->sel_mask + (size_t)w * out->max_sel;
        std::vector<int> idx;
        for (int k = 1; k <= n; ++k) {
            idx.resize((size_t)k);
            for (int j = 0; j < k; ++j) idx[(size_t)j] = j;
            for (;;) {
                uint64_t m = 0;
                for (int j = 0; j < k; ++j) m |= 1ull << idx[(size_t)j];
                if (sel_valid(m, vs->var_pos + v0, vs->var_n_removed + v0, sh.nadd.data() + v0)) {
                    if (k_out >= out->max_sel)
                        return set_err(PLB_ERR_SHAPE, "window %d returns more than max_sel = %d haplotypes", w, out->max_sel);
                    if (out->sel_score) out->sel_score[(size_t)w * out->max_sel + k_out] = nan_v;
                    masks[k_out++] = m;
                }
                int j = k - 1;
                while (j >= 0 && idx[(size_t)j] == n - k + j) --j;
                if (j < 0) break;
                ++idx[(size_t)j];
                for (int q = j + 1; q < k; ++q) idx[(size_t)q] = idx[(size_t)q - 1] + 1;
            }
        }
        out->n_sel[w] = k_out;
    }
    const int Wf_all = (int)filt.size();
    if (Wf_all == 0) return PLB_OK;
    if (out->max_sel < std::min(cap, orig_cap))
        return set_err(PLB_ERR_SHAPE, "max_sel = %d < min(max_haplotypes, original_max_haplotypes) - 1", out->max_sel);

    // ---- processing order: decreasing variant count; two groups (alternate windows of that order) when the batch is
    //      large enough for each group to fill the GPU
    std::stable_sort(filt.begin(), filt.end(), [&](int a, int b) {
        return vs->win_var_off[a + 1] - vs->win_var_off[a] > vs->win_var_off[b + 1] - vs->win_var_off[b];
    });
    int n_groups = Wf_all >= 2048 ? 2 : 1;
    if (const char* e = getenv("PLB_SELECT_GROUPS")) n_groups = std::max(1, std::min({atoi(e), 2, Wf_all}));   // tests
    std::vector<SelPlan> groups((size_t)n_groups);
    for (int i = 0; i < Wf_all; ++i) groups[(size_t)(i % n_groups)].filt.push_back(filt[(size_t)i]);
    const int max_trials = orig_cap + 1;
    double cells_total = 0;

    // ---- per group: the sampled-read batch (variantFilter.pyx:253-277: every sampleRate-th good read), reads as
    //      broken mates, and the variant tables
    auto build_group = [&](int g) -> int {
        SelPlan& P = groups[(size_t)g];
        const int Wf = (int)P.filt.size();
        P.gid = g;
        P.Wf = Wf;
        P.nInd = nInd;
        P.max_rounds = vs->win_var_off[P.filt[0] + 1] - vs->win_var_off[P.filt[0]];
        P.max_trials = max_trials;
        P.ws.resize((size_t)Wf);
        P.we.resize((size_t)Wf);
        P.hs.resize((size_t)Wf);
        P.hoff.resize((size_t)Wf + 1);
        P.zero.assign((size_t)Wf * nInd, 0);
        P.hsoff.resize((size_t)Wf + 1);
        P.slot_off.resize((size_t)Wf * nInd + 1);
        P.voff.resize((size_t)Wf + 1);
        P.win_cells.assign((size_t)Wf, 0);
        // pass 1 (serial, light): sizes -> offsets.  The sampled reads are COPIED into a compact pool of their own
        // (pinned when the scorer provides memory): only every sampleRate-th read is scored, so uploading the caller's
        // whole pool would move several times the bytes the rounds touch.
        std::vector<int> rate((size_t)Wf * nInd, 1);
        std::vector<int64_t> byte_off((size_t)Wf + 1, 0), add_off((size_t)Wf + 1, 0);
        P.hoff[0] = 0;
        P.hsoff[0] = 0;
        P.slot_off[0] = 0;
        P.voff[0] = 0;
        for (int k = 0; k < Wf; ++k) {
            const int w = P.filt[(size_t)k];
            P.hoff[(size_t)k + 1] = k + 1;
            P.hsoff[(size_t)k + 1] = P.hsoff[(size_t)k] + (rb->hap_seq_off[w + 1] - rb->hap_seq_off[w]);
            const int size = rb->win_end[w] - rb->win_start[w];
            int64_t bytes = 0;
            for (int i = 0; i < nInd; ++i) {
                int64_t n_s = 0;
                if (need_reads) {
                    const int64_t wi = (int64_t)w * nInd + i;
                    const int64_t b0 = rb->wi_slot_off[wi];
                    const int n_good = rb->wi_n_good[wi];
                    if (n_good > 0) {
                        const int r_first = rb->slot_read[b0];
                        if (r_first < 0 || r_first >= rb->n_reads)
                            return set_err(PLB_ERR_ARG, "slot %lld: read index out of range", (long long)b0);
                        const int64_t rlen = rb->read_seq_off[r_first + 1] - rb->read_seq_off[r_first];
                        const int64_t mean_cov = rlen * n_good / size;
                        const int rt = (int)std::max<int64_t>(1, mean_cov / so->coverage_sampling_level);
                        rate[(size_t)k * nInd + i] = rt;
                        for (int t = 0; t < n_good; t += rt) {
                            const int r = rb->slot_read[b0 + t];
                            if (r < 0 || r >= rb->n_reads)
                                return set_err(PLB_ERR_ARG, "slot %lld: read index out of range", (long long)(b0 + t));
                            const int64_t L = rb->read_seq_off[r + 1] - rb->read_seq_off[r];
                            if (L < 0 || L > 32767) return set_err(PLB_ERR_SHAPE, "read %d length %lld out of range", r, (long long)L);
                            bytes += L;
                            ++n_s;
                        }
                    }
                }
                P.slot_off[(size_t)k * nInd + i + 1] = P.slot_off[(size_t)k * nInd + i] + n_s;
            }
            byte_off[(size_t)k + 1] = byte_off[(size_t)k] + bytes;
            P.win_cells[(size_t)k] = 16 * bytes;
            cells_total += (double)P.win_cells[(size_t)k];   // the reference-haplotype pass
            const int v0 = vs->win_var_off[w], v1 = vs->win_var_off[w + 1];
            P.voff[(size_t)k + 1] = P.voff[(size_t)k] + (v1 - v0);
            add_off[(size_t)k + 1] = add_off[(size_t)k] + (vs->var_added_off[v1] - vs->var_added_off[v0]);
        }
        const int64_t n_s_all = P.slot_off[(size_t)Wf * nInd], n_bytes = byte_off[(size_t)Wf];
        const int nvar_g = P.voff[(size_t)Wf];
        if (n_s_all > INT32_MAX) return set_err(PLB_ERR_SHAPE, "too many sampled reads in one group");
        P.ref.resize((size_t)P.hsoff[(size_t)Wf]);
        P.slot.resize((size_t)n_s_all);
        P.pos.resize((size_t)nvar_g);
        P.nrem.resize((size_t)nvar_g);
        P.nadd.resize((size_t)nvar_g);
        P.type.resize((size_t)nvar_g);
        P.nsup.resize((size_t)nvar_g);
        P.aoff.resize((size_t)nvar_g + 1);
        P.add.resize((size_t)add_off[(size_t)Wf]);
        P.aoff[0] = 0;
        int64_t* c_off = nullptr;
        uint8_t *c_seq = nullptr, *c_qual = nullptr, *c_mapq = nullptr, *c_qc = nullptr;
        int32_t *c_pos = nullptr, *c_end = nullptr;
        if (need_reads) {
            c_off = (int64_t*)alloc(((size_t)n_s_all + 1) * 8);
            c_seq = (uint8_t*)alloc((size_t)n_bytes + 64);
            c_qual = (uint8_t*)alloc((size_t)n_bytes + 64);
            c_pos = (int32_t*)alloc((size_t)n_s_all * 4 + 4);
            c_end = (int32_t*)alloc((size_t)n_s_all * 4 + 4);
            c_mapq = (uint8_t*)alloc((size_t)n_s_all + 4);
            c_qc = (uint8_t*)alloc((size_t)n_s_all + 4);
            if (!c_off || !c_seq || !c_qual || !c_pos || !c_end || !c_mapq || !c_qc)
                return set_err(PLB_ERR_NOMEM, "host allocation for the sampled reads failed");
            c_off[0] = 0;
        }
        // pass 2 (parallel): copies
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (Wf > 256)
        for (int k = 0; k < Wf; ++k) {
            const int w = P.filt[(size_t)k];
            P.ws[(size_t)k] = rb->win_start[w];
            P.we[(size_t)k] = rb->win_end[w];
            P.hs[(size_t)k] = rb->hap_start[w];
            memcpy(P.ref.data() + P.hsoff[(size_t)k], rb->hap_seq + rb->hap_seq_off[w], (size_t)(P.hsoff[(size_t)k + 1] - P.hsoff[(size_t)k]));
            int64_t bo = byte_off[(size_t)k];
            for (int i = 0; i < nInd && need_reads; ++i) {
                const int64_t wi = (int64_t)w * nInd + i;
                const int64_t b0 = rb->wi_slot_off[wi];
                const int n_good = rb->wi_n_good[wi], rt = rate[(size_t)k * nInd + i];
                int64_t sl = P.slot_off[(size_t)k * nInd + i];
                for (int t = 0; t < n_good; t += rt, ++sl) {
                    const int r = rb->slot_read[b0 + t];
                    const int64_t o = rb->read_seq_off[r], L = rb->read_seq_off[r + 1] - o;
                    memcpy(c_seq + bo, rb->read_seq + o, (size_t)L);
                    memcpy(c_qual + bo, rb->read_qual + o, (size_t)L);
                    bo += L;
                    c_off[sl + 1] = bo;
                    c_pos[sl] = rb->read_pos[r];
                    c_end[sl] = rb->read_end[r];
                    c_mapq[sl] = rb->read_mapq[r];
                    c_qc[sl] = rb->read_qcfail[r];
                    P.slot[(size_t)sl] = (int32_t)sl;
                }
            }
            const int v0 = vs->win_var_off[w], nv = vs->win_var_off[w + 1] - v0;
            const int q0 = P.voff[(size_t)k];
            const int64_t a0 = add_off[(size_t)k], src0 = vs->var_added_off[v0];
            for (int j = 0; j < nv; ++j) {
                P.pos[(size_t)q0 + j] = vs->var_pos[v0 + j];
                P.nrem[(size_t)q0 + j] = vs->var_n_removed[v0 + j];
                P.nadd[(size_t)q0 + j] = sh.nadd[(size_t)v0 + j];
                P.type[(size_t)q0 + j] = sh.type[(size_t)v0 + j];
                P.nsup[(size_t)q0 + j] = vs->var_n_support[v0 + j];
                P.aoff[(size_t)q0 + j + 1] = a0 + (vs->var_added_off[v0 + j + 1] - src0);
            }
            if (add_off[(size_t)k + 1] > a0) memcpy(P.add.data() + a0, vs->var_added + src0, (size_t)(add_off[(size_t)k + 1] - a0));
        }
        PlbWindowBatch& hsb = P.batch;
        hsb = *rb;
        if (need_reads) {   // the compact pool of sampled reads
            hsb.n_reads = (int32_t)n_s_all;
            hsb.read_seq_off = c_off;
            hsb.read_seq = c_seq;
            hsb.read_qual = c_qual;
            hsb.read_pos = c_pos;
            hsb.read_end = c_end;
            hsb.read_mapq = c_mapq;
            hsb.read_qcfail = c_qc;
        }
        hsb.n_windows = Wf;
        hsb.n_haps = Wf;
        hsb.n_slots = (int64_t)P.slot.size();
        hsb.win_hap_off = P.hoff.data();
        hsb.win_start = P.ws.data();
        hsb.win_end = P.we.data();
        hsb.hap_start = P.hs.data();
        hsb.hap_seq_off = P.hsoff.data();
        hsb.hap_seq = P.ref.data();
        hsb.wi_slot_off = P.slot_off.data();
        hsb.wi_n_good = P.zero.data();
        hsb.wi_n_bad = P.zero.data();
        hsb.slot_read = P.slot.data();
        hsb.max_variants = 0;
        hsb.win_n_var = nullptr;
        hsb.hap_var_mask = nullptr;
        hsb.var_prior = nullptr;
        if (need_reads) {
            if (hsb.n_slots > 0 && (!rb->slot_read || !rb->read_seq_off || !rb->read_seq || !rb->read_qual || !rb->read_pos ||
                                    !rb->read_end || !rb->read_mapq || !rb->read_qcfail))
                return set_err(PLB_ERR_ARG, "NULL read array in batch");
            if (opt && opt->calc_flank_score)
                for (int k = 0; k < Wf; ++k)
                    if (P.ws[(size_t)k] - P.hs[(size_t)k] <= 0)
                        return set_err(PLB_ERR_ARG, "window %d: calc_flank_score needs a positive flank", P.filt[(size_t)k]);
        }
        P.state.resize((size_t)Wf);
        for (int k = 0; k < Wf; ++k) {
            const int n = P.voff[(size_t)k + 1] - P.voff[(size_t)k];
            SelEntryState& s = P.state[(size_t)k];
            s.order.resize((size_t)n);
            for (int j = 0; j < n; ++j) s.order[(size_t)j] = j;
            const int32_t* ns = P.nsup.data() + P.voff[(size_t)k];
            std::stable_sort(s.order.begin(), s.order.end(), [&](int a, int b) { return ns[a] > ns[b]; });
            s.heap.reserve((size_t)orig_cap + 1);
        }
        P.trial_mask.resize((size_t)Wf * max_trials);
        P.n_trials.resize((size_t)Wf);
        int rc_ = prepare(P);
        if (rc_) return rc_;
        if (!P.r_hoff) {   // the scorer did not supply (pinned) memory for the per-round arrays
            P.own_hoff.resize((size_t)P.Wf + 1);
            P.own_hsoff.resize((size_t)P.Wf * max_trials + 1);
            P.own_mask.resize((size_t)P.Wf * max_trials);
            P.own_scores.resize((size_t)P.Wf * max_trials);
            P.r_hoff = P.own_hoff.data();
            P.r_hsoff = P.own_hsoff.data();
            P.mask_c = P.own_mask.data();
            P.scores = P.own_scores.data();
        }
        return PLB_OK;
    };

    double t_host = 0, n_trials_total = 0, tb_trials = 0, tb_len = 0, tb_heap = 0;
    int rounds = 0;
    // While the heap is not full nothing is evicted, so the trial SETS of the first rounds do not depend on any score:
    // round r tries v_r alone and v_r joined to every earlier (valid) trial.  The first K rounds, K = the largest with
    // 2^K - 1 <= originalMaxHaplotypes - 1 (5 for the default 49: at most 31 sets), are therefore scored in ONE launch;
    // their heap pushes are replayed afterwards in the reference's order (which does depend on the scores, through
    // sorted(heap)).  That removes the latency-bound launches with 1, 2, 4 ... trials per window.
    int prologue = 0;
    while (prologue < 6 && (1 << (prologue + 1)) - 1 <= orig_cap) ++prologue;
    if (getenv("PLB_SELECT_NO_PROLOGUE")) prologue = 0;   // tests: the plain round-by-round schedule
    // host phase 1 of a group's next round: the trial haplotypes (variantFilter.pyx:452-476: {v_r}, then v_r joined to
    // every kept set in sorted(heap) order) and their lengths.  Returns 1 when there is a round to score.
    auto begin_round = [&](SelPlan& P) -> int {
        const double th0 = now_ms();
        const int r = P.r;
        int Wr = 0;
        if (r < P.max_rounds)
            while (Wr < P.Wf && P.voff[(size_t)Wr + 1] - P.voff[(size_t)Wr] > r) ++Wr;
        P.Wr = Wr;
        if (Wr == 0) return 0;
        P.in_prologue = r == 0 && prologue > 1;
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (Wr > 256)
        for (int k = 0; k < Wr; ++k) {
            SelEntryState& s = P.state[(size_t)k];
            const int v0 = P.voff[(size_t)k];
            const SelKeys keys{P.pos.data() + v0, P.type.data() + v0, P.nrem.data() + v0};
            uint64_t* m = P.trial_mask.data() + (size_t)k * max_trials;
            if (P.in_prologue) {   // every trial set of rounds 0 .. min(prologue, nVar) - 1
                const int nr = std::min(prologue, P.voff[(size_t)k + 1] - v0);
                int n = 0;
                for (int q = 0; q < nr; ++q) {
                    const uint64_t bit = 1ull << s.order[(size_t)q];
                    const int n_prev = n;
                    m[n++] = bit;
                    for (int j = 0; j < n_prev; ++j) {
                        const uint64_t both = m[j] | bit;
                        if (sel_valid(both, keys.pos, keys.nrem, P.nadd.data() + v0)) m[n++] = both;
                    }
                }
                P.n_trials[(size_t)k] = n;
                continue;
            }
            const uint64_t bit = 1ull << s.order[(size_t)r];
            int n = 0;
            m[n++] = bit;
            std::vector<SelEntry> old = s.heap;
            py_sort(old, keys, false);
            for (const SelEntry& e : old) {
                const uint64_t both = e.mask | bit;
                if (sel_valid(both, keys.pos, keys.nrem, P.nadd.data() + v0)) m[n++] = both;
            }
            P.n_trials[(size_t)k] = n;
        }
        const double th_a = now_ms();
        tb_trials += th_a - th0;
        P.r_hoff[0] = 0;
        for (int k = 0; k < Wr; ++k) P.r_hoff[(size_t)k + 1] = P.r_hoff[(size_t)k] + P.n_trials[(size_t)k];
        const int nh = P.r_hoff[(size_t)Wr];
        P.nh = nh;
        // the masks in haplotype order, and the sequence lengths
        int bad_shape = 0;
        P.r_hsoff[0] = 0;
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (Wr > 256)
        for (int k = 0; k < Wr; ++k) {
            const int ws = P.ws[(size_t)k], we = P.we[(size_t)k];
            const int left = std::min(ws - P.hs[(size_t)k], ws);
            int64_t min_hap = INT64_MAX;
            const int v0 = P.voff[(size_t)k];
            for (int j = 0; j < P.n_trials[(size_t)k]; ++j) {
                const uint64_t mk = P.trial_mask[(size_t)k * max_trials + j];
                P.mask_c[(size_t)P.r_hoff[(size_t)k] + j] = mk;
                int64_t len = fast_hap_length(mk, P.hsoff[(size_t)k + 1] - P.hsoff[(size_t)k], we, P.pos.data() + v0,
                                              P.nrem.data() + v0, P.nadd.data() + v0);
                if (len < 0 || check_len) {
                    const int64_t fast = len;
                    len = 0;
                    walk_haplotype(ws, we, ws - left, (int)(P.hsoff[(size_t)k + 1] - P.hsoff[(size_t)k]), mk, v0, P.pos.data(),
                                   P.nrem.data(), P.aoff.data(), [&](int, int, int n) { len += n; });
                    if (fast >= 0 && fast != len) bad_shape = 3;   // PLB_SELECT_CHECK=1: the shortcut must agree with the walk
                }
                P.r_hsoff[(size_t)P.r_hoff[(size_t)k] + j + 1] = len;
                if (len > PLB_MAX_HAP_LEN) bad_shape = 1;
                min_hap = std::min(min_hap, len);
            }
            // every scored read needs hapLen >= readLen + 15 (calign.pyx:256-259)
            for (int64_t sl = P.slot_off[(size_t)k * nInd]; need_reads && sl < P.slot_off[(size_t)(k + 1) * nInd]; ++sl) {
                const int64_t rl = P.batch.read_seq_off[sl + 1] - P.batch.read_seq_off[sl];
                if (rl >= PLB_KMER && rl + 15 > min_hap) bad_shape = 2;
            }
        }
        if (bad_shape == 3) return set_err(PLB_ERR_ARG, "internal: fast haplotype length disagrees with the walk");
        if (bad_shape)
            return set_err(PLB_ERR_SHAPE, bad_shape == 1 ? "a trial haplotype is longer than 16384 (chaplotype.pyx:180-183)"
                                                         : "a trial haplotype is shorter than readLen + 15 (calign.pyx:256-259)");
        // lengths -> offsets: per-window totals, a prefix over the windows, then every window's own running sum
        {
            std::vector<int64_t> lens_check;
            if (check_len) lens_check.assign(P.r_hsoff + 1, P.r_hsoff + 1 + nh);
            std::vector<int64_t>& base = P.win_base;
            base.resize((size_t)Wr + 1);
            base[0] = 0;
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (Wr > 256)
            for (int k = 0; k < Wr; ++k) {
                int64_t sum = 0;
                for (int h = P.r_hoff[(size_t)k]; h < P.r_hoff[(size_t)k + 1]; ++h) sum += P.r_hsoff[(size_t)h + 1];
                base[(size_t)k + 1] = sum;
            }
            for (int k = 0; k < Wr; ++k) base[(size_t)k + 1] += base[(size_t)k];
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (Wr > 256)
            for (int k = 0; k < Wr; ++k) {
                int64_t run = base[(size_t)k];
                for (int h = P.r_hoff[(size_t)k]; h < P.r_hoff[(size_t)k + 1]; ++h) {
                    run += P.r_hsoff[(size_t)h + 1];
                    P.r_hsoff[(size_t)h + 1] = run;
                }
            }
            if (check_len) {   // PLB_SELECT_CHECK=1: offsets must be the running sum of the lengths, window by window
                int64_t run = 0;
                for (int h = 0; h < nh; ++h) {
                    run += lens_check[(size_t)h];
                    if (P.r_hsoff[(size_t)h + 1] != run) return set_err(PLB_ERR_ARG, "internal: haplotype offsets inconsistent");
                }
                if (P.r_hsoff[0] != 0) return set_err(PLB_ERR_ARG, "internal: haplotype offsets inconsistent");
            }
        }
        const double th_b = now_ms();
        tb_len += th_b - th_a;
        t_host += th_b - th0;
        return 1;
    };
    // host phase 2: heap updates in the reference's order (:459-486)
    auto end_round = [&](SelPlan& P) {
        const double th1 = now_ms();
        const int Wr = P.Wr;
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (Wr > 256)
        for (int k = 0; k < Wr; ++k) {
            SelEntryState& s = P.state[(size_t)k];
            const int v0 = P.voff[(size_t)k];
            const SelKeys keys{P.pos.data() + v0, P.type.data() + v0, P.nrem.data() + v0};
            if (P.in_prologue) {   // replay rounds 0 .. nr-1 with the scores of the one launch
                const int nr = std::min(prologue, P.voff[(size_t)k + 1] - v0);
                const int nt = P.n_trials[(size_t)k];
                const uint64_t* tm = P.mask_c + (size_t)P.r_hoff[(size_t)k];
                const double* ts = P.scores + (size_t)P.r_hoff[(size_t)k];
                auto score_of = [&](uint64_t mk) {
                    for (int j = 0; j < nt; ++j)
                        if (tm[j] == mk) return ts[j];
                    return -1e20;   // unreachable: every set the replay forms was enumerated
                };
                std::vector<SelEntry> old;
                for (int q = 0; q < nr; ++q) {
                    const uint64_t bit = 1ull << s.order[(size_t)q];
                    old = s.heap;
                    py_sort(old, keys, false);
                    auto push = [&](uint64_t mk) {
                        const SelEntry e{score_of(mk), mk};
                        if ((int)s.heap.size() < orig_cap)
                            heap_push(s.heap, e, keys);
                        else
                            heap_pushpop(s.heap, e, keys);
                    };
                    push(bit);
                    for (const SelEntry& e : old) {
                        const uint64_t both = e.mask | bit;
                        if (sel_valid(both, keys.pos, keys.nrem, P.nadd.data() + v0)) push(both);
                    }
                }
                s.n_done += nt;
                continue;
            }
            for (int j = 0; j < P.n_trials[(size_t)k]; ++j) {
                const SelEntry e{P.scores[(size_t)P.r_hoff[(size_t)k] + j], P.mask_c[(size_t)P.r_hoff[(size_t)k] + j]};
                if ((int)s.heap.size() < orig_cap)
                    heap_push(s.heap, e, keys);
                else
                    heap_pushpop(s.heap, e, keys);
            }
            s.n_done += P.n_trials[(size_t)k];
        }
        n_trials_total += P.nh;
        for (int k = 0; k < Wr; ++k) cells_total += (double)P.win_cells[(size_t)k] * P.n_trials[(size_t)k];
        P.r = P.in_prologue ? prologue : P.r + 1;
        P.in_prologue = false;
        const double th2 = now_ms();
        tb_heap += th2 - th1;
        t_host += th2 - th1;
    };
    auto start = [&](SelPlan& P) -> int {   // begin the group's next round and hand it to the scorer
        const int b = begin_round(P);
        if (b < 0) return b;
        P.active = b == 1;
        if (!P.active) return PLB_OK;
        const SelRound R{P.Wr, P.nh, P.r_hoff, P.r_hsoff, P.mask_c, P.n_trials.data()};
        return submit(P, R);
    };
    // group by group: the first group's launch is on the GPU while the second group's plan and read pool are built
    for (int g = 0; g < n_groups; ++g)
        if ((rc = build_group(g)) || (rc = start(groups[(size_t)g]))) return rc;
    for (bool any = true; any;) {
        any = false;
        for (auto& P : groups) {
            if (!P.active) continue;
            if ((rc = wait(P))) return rc;
            end_round(P);
            ++rounds;   // group-rounds: one launch sequence each
            if ((rc = start(P))) return rc;
            any = any || P.active;
        }
        for (auto& P : groups) any = any || P.active;
    }
    // ---- the best max_haplotypes - 1 sets, best first (:497-503)
    const double th2 = now_ms();
    for (auto& P : groups) {
        const int Wf = P.Wf;
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (Wf > 256)
        for (int k = 0; k < Wf; ++k) {
            const int w = P.filt[(size_t)k];
            SelEntryState& s = P.state[(size_t)k];
            const int v0 = P.voff[(size_t)k];
            const SelKeys keys{P.pos.data() + v0, P.type.data() + v0, P.nrem.data() + v0};
            std::vector<SelEntry> fin = s.heap;
            py_sort(fin, keys, true);
            const int n = std::min((int)fin.size(), cap);
            for (int j = 0; j < n; ++j) {
                out->sel_mask[(size_t)w * out->max_sel + j] = fin[(size_t)j].mask;
                if (out->sel_score) out->sel_score[(size_t)w * out->max_sel + j] = fin[(size_t)j].score;
            }
            out->n_sel[w] = n;
            if (out->n_scored) out->n_scored[w] = s.n_done;
        }
    }
    t_host += now_ms() - th2;
    if (getenv("PLB_TRACE"))
        fprintf(stderr, "[plb select] bookkeeping: trial sets %.2f, lengths %.2f, heap updates %.2f, final ranking %.2f ms\n", tb_trials,
                tb_len, tb_heap, now_ms() - th2);
    g_sel_stats.v[4] = t_host;
    g_sel_stats.v[5] = rounds;
    g_sel_stats.v[6] = n_trials_total;
    g_sel_stats.v[8] = cells_total;
    g_sel_stats.v[9] = Wf_all;
    return PLB_OK;
}

extern "C" int plb_select_replay_host(const PlbWindowBatch* rb, const PlbVariantSet* vs, const PlbSelectOptions* so,
                                      plb_trial_score_fn score, void* user, PlbSelectOut* out) {
    if (!score) return set_err(PLB_ERR_ARG, "NULL argument");
    std::vector<int32_t> hw;
    return select_core(
        rb, vs, so, nullptr, out, false, [](size_t) -> void* { return nullptr; }, [](SelPlan&) { return PLB_OK; },
        [&](SelPlan& P, const SelRound& R) -> int {
            hw.resize((size_t)R.nh);
            for (int k = 0; k < R.Wr; ++k)
                for (int h = R.hap_off[k]; h < R.hap_off[k + 1]; ++h) hw[(size_t)h] = P.filt[(size_t)k];
            const int rc = score(user, R.nh, hw.data(), R.mask, P.scores);
            return rc ? set_err(PLB_ERR_ARG, "trial score callback failed (%d)", rc) : PLB_OK;
        },
        [](SelPlan&) { return PLB_OK; });
}

extern "C" int plb_select_haplotypes_host(PlbContext* c, const PlbWindowBatch* rb, const PlbVariantSet* vs,
                                          const PlbSelectOptions* so, const PlbOptions* opt_in, PlbSelectOut* out) {
    if (!c) return set_err(PLB_ERR_ARG, "NULL argument");
    int rc = check_options(opt_in);
    if (rc) return rc;
    if ((rc = require_idle(c, "plb_select_haplotypes_host"))) return rc;
    if (rb && (rb->seq_format != PLB_SEQ_ASCII || rb->qual_bits != 0))
        return set_err(PLB_ERR_UNSUPPORTED, "plb_select_haplotypes_host samples the reads on the host: it takes ASCII batches only");
    PlbOptions opt = *opt_in;
    opt.use_mapq_cap = 0;   // alignSingleRead(read, False), variantFilter.pyx:274-275
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    struct GroupDev {   // device side of one group
        PlbDeviceBatch* base = nullptr;   // reference haplotypes + reads / slots, resident for all rounds
        PlbDeviceBatch* rd = nullptr;     // the round in flight
        Block VB{nullptr, 0};
        cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        double* d_llref = nullptr;
        uint64_t* d_mask = nullptr;
        double* d_score = nullptr;
        SelVars sv{};
    };
    GroupDev gd[2];
    const bool modes = opt.calc_flank_score != 0;
    double t_build = 0, t_score = 0, t_reduce = 0, t_refpass = 0, n_pairs_total = 0;
    double tr_prepare = 0, tr_setup = 0, tr_plan = 0, tr_wait = 0;   // PLB_TRACE: host milliseconds by phase
    auto cleanup = [&]() {
        cudaStreamSynchronize(st);
        for (auto& G : gd) {
            if (G.rd) plb_batch_free(c, G.rd);
            if (G.base) plb_batch_free(c, G.base);
            G.rd = G.base = nullptr;
            if (G.VB.p) block_put(c, G.VB);
            G.VB.p = nullptr;
            for (auto& e : G.ev) {
                if (e) cudaEventDestroy(e);
                e = nullptr;
            }
        }
    };
#define SEL_TRY(expr)        \
    do {                     \
        int rc_ = (expr);    \
        if (rc_) return rc_; \
    } while (0)
#define SEL_CU(call)                                                                                                  \
    do {                                                                                                              \
        cudaError_t e_ = (call);                                                                                      \
        if (e_ != cudaSuccess)                                                                                        \
            return set_err(PLB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
    // reference-haplotype pass of a group: uploads its reads / slots once and leaves ll_ref[slot] on the device
    auto prepare = [&](SelPlan& P) -> int {
        const double tp0 = now_ms();
        GroupDev& G = gd[P.gid];
        const int Wf = P.Wf, nvar = (int)P.pos.size();
        SEL_TRY(prepare_batch(c, &P.batch, &G.base, nullptr, false));
        Layout L;
        const size_t o_llref = L.take((size_t)P.batch.n_slots * 8 + 64), o_voff = L.take((size_t)(Wf + 1) * 4),
                     o_vpos = L.take((size_t)nvar * 4), o_vnr = L.take((size_t)nvar * 4),
                     o_vaoff = L.take((size_t)(nvar + 1) * 8), o_vadd = L.take(P.add.size() + 64),
                     o_mask = L.take((size_t)Wf * P.max_trials * 8), o_score = L.take((size_t)Wf * P.max_trials * 8);
        SEL_TRY(block_get(c, L.off + 256, &G.VB));
        for (auto& e : G.ev) SEL_CU(cudaEventCreate(&e));
        G.d_llref = at<double>(G.VB, o_llref);
        G.d_mask = at<uint64_t>(G.VB, o_mask);
        G.d_score = at<double>(G.VB, o_score);
        // per-round host arrays in pinned memory: the copies must not block the host while the other group runs
        const size_t nt = (size_t)Wf * P.max_trials;
        P.r_hoff = (int32_t*)pin_alloc(c, ((size_t)Wf + 1) * 4);
        P.r_hsoff = (int64_t*)pin_alloc(c, (nt + 1) * 8);
        P.mask_c = (uint64_t*)pin_alloc(c, nt * 8);
        P.scores = (double*)pin_alloc(c, nt * 8);
        if (!P.r_hoff || !P.r_hsoff || !P.mask_c || !P.scores) return set_err(PLB_ERR_NOMEM, "pinned host allocation failed");
        SEL_CU(cudaMemcpyAsync(at<uint8_t>(G.VB, o_voff), P.voff.data(), (size_t)(Wf + 1) * 4, cudaMemcpyHostToDevice, st));
        if (nvar) {
            SEL_CU(cudaMemcpyAsync(at<uint8_t>(G.VB, o_vpos), P.pos.data(), (size_t)nvar * 4, cudaMemcpyHostToDevice, st));
            SEL_CU(cudaMemcpyAsync(at<uint8_t>(G.VB, o_vnr), P.nrem.data(), (size_t)nvar * 4, cudaMemcpyHostToDevice, st));
            if (!P.add.empty())
                SEL_CU(cudaMemcpyAsync(at<uint8_t>(G.VB, o_vadd), P.add.data(), P.add.size(), cudaMemcpyHostToDevice, st));
        }
        SEL_CU(cudaMemcpyAsync(at<uint8_t>(G.VB, o_vaoff), P.aoff.data(), (size_t)(nvar + 1) * 8, cudaMemcpyHostToDevice, st));
        G.sv = SelVars{at<int32_t>(G.VB, o_voff), at<int32_t>(G.VB, o_vpos), at<int32_t>(G.VB, o_vnr), at<int64_t>(G.VB, o_vaoff),
                       at<uint8_t>(G.VB, o_vadd)};
        if (P.gid == 0) SEL_CU(cudaMemsetAsync(c->d_ctr(), 0, sizeof(Counters), st));
        SEL_TRY(copy_seq_for_windows(c, G.base, &P.batch, 0, Wf, st));
        SEL_TRY(plan_chunk(c, G.base, &P.batch, 0, Wf, st));
        SEL_TRY(derive_all(c, G.base, st));
        if (modes) SEL_TRY(mode_queues(c, G.base, false));
        PlbLoglikOut llo{nullptr, G.d_llref, nullptr};
        SEL_CU(cudaEventRecord(G.ev[0], st));
        SEL_TRY(launch_windows(c, G.base, G.base->chunks[0], &opt, nullptr, &llo, st, false, modes ? G.base->mq : G.base->q));
        SEL_CU(cudaEventRecord(G.ev[1], st));
        n_pairs_total += (double)G.base->d.n_pairs;
        tr_prepare += now_ms() - tp0;
        return PLB_OK;
    };
    // one round of a group on the device: build the trial haplotypes, score every sampled read against them, reduce,
    // send the scores back; asynchronous - `wait` below completes it
    auto submit = [&](SelPlan& P, const SelRound& R) -> int {
        GroupDev& G = gd[P.gid];
        PlbWindowBatch hr = P.batch;
        hr.n_windows = R.Wr;
        hr.n_haps = R.nh;
        hr.n_slots = P.slot_off[(size_t)R.Wr * P.nInd];
        hr.win_hap_off = R.hap_off;
        hr.hap_seq_off = R.hap_seq_off;
        hr.hap_seq = nullptr;   // built on the device
        const double ts0 = now_ms();
        SEL_TRY(prepare_batch(c, &hr, &G.rd, G.base));
        SEL_TRY(copy_meta(c, G.rd, &hr, 0, R.Wr, 0, -1, st));
        SEL_CU(cudaMemcpyAsync(G.d_mask, R.mask, (size_t)R.nh * 8, cudaMemcpyHostToDevice, st));
        const double ts1 = now_ms();
        SEL_TRY(plan_chunk(c, G.rd, &hr, 0, R.Wr, st));
        const double ts2 = now_ms();
        tr_setup += ts1 - ts0;
        tr_plan += ts2 - ts1;
        SEL_TRY(derive_all(c, G.rd, st));
        if (modes) SEL_TRY(mode_queues(c, G.rd, false));
        SEL_CU(cudaEventRecord(G.ev[2], st));
        const int grid = std::max(1, std::min((R.nh + kBuildWarps - 1) / kBuildWarps, c->n_sm * 16));
        k_build_haps<<<grid, 32 * kBuildWarps, 0, st>>>(R.nh, G.rd->d.hap_win, G.d_mask, G.base->d.hap_seq_off, G.base->d.hap_seq,
                                                       G.rd->d.win_start, G.rd->d.win_end, G.rd->d.hap_start, G.sv,
                                                       G.rd->d.hap_seq_off, (uint8_t*)G.rd->d.hap_seq);
        SEL_TRY(launch_check(c, "k_build_haps"));
        SEL_CU(cudaEventRecord(G.ev[3], st));
        SEL_TRY(launch_windows(c, G.rd, G.rd->chunks[0], &opt, nullptr, nullptr, st, c->timing, modes ? G.rd->mq : G.rd->q));
        if (c->timing) {   // one ring entry per round launch (plb_kernel_times averages them)
            c->kev_chunks[c->n_timed % kTimingRing] = 1;
            c->n_timed++;
        }
        SEL_CU(cudaEventRecord(G.ev[4], st));
        k_trial_score<<<(R.nh + 3) / 4, 128, 0, st>>>(G.rd->d, G.rd->ll_scratch, G.d_llref, R.nh, G.d_score);
        SEL_TRY(launch_check(c, "k_trial_score"));
        SEL_CU(cudaEventRecord(G.ev[5], st));
        SEL_CU(cudaMemcpyAsync(P.scores, G.d_score, (size_t)R.nh * 8, cudaMemcpyDeviceToHost, st));
        SEL_CU(cudaEventRecord(G.ev[6], st));
        tr_wait += now_ms() - ts2;
        return PLB_OK;
    };
    auto wait = [&](SelPlan& P) -> int {
        GroupDev& G = gd[P.gid];
        const double tw0 = now_ms();
        SEL_CU(cudaEventSynchronize(G.ev[6]));
        float ms = 0;
        cudaEventElapsedTime(&ms, G.ev[2], G.ev[3]);
        t_build += ms;
        cudaEventElapsedTime(&ms, G.ev[3], G.ev[4]);
        t_score += ms;
        cudaEventElapsedTime(&ms, G.ev[4], G.ev[5]);
        t_reduce += ms;
        if (P.r == 0) {
            cudaEventElapsedTime(&ms, G.ev[0], G.ev[1]);
            t_refpass += ms;
        }
        n_pairs_total += (double)G.rd->d.n_pairs;
        batch_release_done(c, G.rd);
        G.rd = nullptr;
        tr_wait += now_ms() - tw0;
        return PLB_OK;
    };
#undef SEL_TRY
#undef SEL_CU
    static const bool trace = getenv("PLB_TRACE") != nullptr;
    const double t_call0 = now_ms();
    pin_reset(c);   // one arena for the whole call: sampled reads, per-round arrays, tile lists
    rc = select_core(rb, vs, so, &opt, out, true, [&](size_t bytes) { return pin_alloc(c, bytes); }, prepare, submit, wait);
    if (trace)
        fprintf(stderr, "[plb select] total %.2f ms: prepare %.2f | rounds: batch setup %.2f, plan %.2f, launch+wait %.2f | host bookkeeping %.2f\n",
                now_ms() - t_call0, tr_prepare, tr_setup, tr_plan, tr_wait, g_sel_stats.v[4]);
    cleanup();
    if (rc) return rc;
    g_sel_stats.v[0] = t_refpass;
    g_sel_stats.v[1] = t_build;
    g_sel_stats.v[2] = t_score;
    g_sel_stats.v[3] = t_reduce;
    g_sel_stats.v[7] = n_pairs_total;
    return PLB_OK;
}

extern "C" int plb_best_score_haplotypes_host(PlbContext* c, const PlbWindowBatch* hb, const PlbOptions* opt_in,
                                              double* score_out) {
    if (!c || !hb || !score_out) return set_err(PLB_ERR_ARG, "NULL argument");
    int rc = check_options(opt_in);
    if (rc) return rc;
    PlbOptions opt = *opt_in;
    opt.use_mapq_cap = 0;   // alignSingleRead(read, False)
    const int W = hb->n_windows, nInd = hb->n_individuals;
    if (W <= 0 || hb->n_haps <= 0) return PLB_OK;
    if ((rc = plb_validate(hb, &opt, 0))) return rc;
    CU(cudaSetDevice(c->device));
    // the good reads of every (window, individual), as broken mates: a bare alignReadToHaplotype per read
    const int64_t nwi = (int64_t)W * nInd;
    std::vector<int64_t> slot_off((size_t)nwi + 1, 0);
    for (int64_t wi = 0; wi < nwi; ++wi) slot_off[(size_t)wi + 1] = slot_off[(size_t)wi] + hb->wi_n_good[wi];
    std::vector<int32_t> slots((size_t)slot_off[(size_t)nwi]), zero((size_t)nwi, 0);
    for (int64_t wi = 0; wi < nwi; ++wi)
        for (int t = 0; t < hb->wi_n_good[wi]; ++t) slots[(size_t)slot_off[(size_t)wi] + t] = hb->slot_read[hb->wi_slot_off[wi] + t];
    PlbWindowBatch g = *hb;
    g.n_slots = slot_off[(size_t)nwi];
    g.wi_slot_off = slot_off.data();
    g.slot_read = slots.data();
    g.wi_n_good = zero.data();
    g.wi_n_bad = zero.data();
    g.max_variants = 0;
    g.win_n_var = nullptr;
    g.hap_var_mask = nullptr;
    g.var_prior = nullptr;
    PlbDeviceBatch* db = nullptr;
    if ((rc = plb_batch_upload(c, &g, &db))) return rc;
    cudaStream_t st = c->stream;
    Block out{nullptr, 0};
    const bool modes = opt.calc_flank_score != 0;
    rc = block_get(c, (size_t)g.n_haps * 8 + 256, &out);
    if (!rc && modes) rc = mode_queues(c, db, false);
    if (!rc) rc = launch_windows(c, db, db->chunks[0], &opt, nullptr, nullptr, st, false, modes ? db->mq : db->q);
    if (!rc) {
        k_hap_score<<<(g.n_haps + 3) / 4, 128, 0, st>>>(db->d, db->ll_scratch, g.n_haps, (double*)out.p);
        rc = launch_check(c, "k_hap_score");
    }
    cudaError_t e = cudaSuccess;
    if (!rc) e = cudaMemcpyAsync(score_out, out.p, (size_t)g.n_haps * 8, cudaMemcpyDeviceToHost, st);
    const cudaError_t e2 = cudaStreamSynchronize(st);
    if (out.p) block_put(c, out);
    plb_batch_free(c, db);
    if (rc) return rc;
    if (e != cudaSuccess || e2 != cudaSuccess)
        return set_err(PLB_ERR_CUDA, "plb_best_score_haplotypes_host: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return PLB_OK;
}

extern "C" int plb_best_score_genotypes_host(PlbContext* c, const PlbWindowBatch* hb, const PlbOptions* opt_in,
                                             int32_t target_coverage, int32_t n_pairs, const int32_t* hap1,
                                             const int32_t* hap2, double* score_out) {
    if (!c || !hb || !score_out || (n_pairs > 0 && (!hap1 || !hap2))) return set_err(PLB_ERR_ARG, "NULL argument");
    if (target_coverage <= 0) return set_err(PLB_ERR_ARG, "target_coverage must be positive (variantFilter.pyx:251)");
    int rc = check_options(opt_in);
    if (rc) return rc;
    if ((rc = require_idle(c, "plb_best_score_genotypes_host"))) return rc;
    PlbOptions opt = *opt_in;
    opt.use_mapq_cap = 0;   // alignSingleRead(read, False)
    const int W = hb->n_windows, nInd = hb->n_individuals;
    if (n_pairs <= 0) return PLB_OK;
    if (W <= 0 || hb->n_haps <= 0) return set_err(PLB_ERR_ARG, "pairs given for an empty batch");
    if (hb->seq_format != PLB_SEQ_ASCII || hb->qual_bits != 0)
        return set_err(PLB_ERR_UNSUPPORTED, "plb_best_score_genotypes_host takes byte-per-base batches");
    if ((rc = plb_validate(hb, &opt, 0))) return rc;
    // a pair names two haplotypes of ONE window
    std::vector<int32_t> hap_win((size_t)hb->n_haps);
    for (int w = 0; w < W; ++w)
        for (int h = hb->win_hap_off[w]; h < hb->win_hap_off[w + 1]; ++h) hap_win[(size_t)h] = w;
    for (int p = 0; p < n_pairs; ++p) {
        if (hap1[p] < 0 || hap1[p] >= hb->n_haps || hap2[p] < 0 || hap2[p] >= hb->n_haps)
            return set_err(PLB_ERR_ARG, "pair %d: haplotype index out of range", p);
        if (hap_win[(size_t)hap1[p]] != hap_win[(size_t)hap2[p]])
            return set_err(PLB_ERR_ARG, "pair %d: the two haplotypes belong to different windows", p);
    }
    CU(cudaSetDevice(c->device));
    // the sampled good reads of every (window, individual) (variantFilter.pyx:253-277: every sampleRate-th one), as
    // broken mates: a bare alignReadToHaplotype per read
    const int64_t nwi = (int64_t)W * nInd;
    std::vector<int64_t> slot_off((size_t)nwi + 1, 0);
    std::vector<int32_t> slots, zero((size_t)nwi, 0);
    for (int w = 0; w < W; ++w) {
        const int64_t size = (int64_t)hb->win_end[w] - hb->win_start[w];
        if (size <= 0) return set_err(PLB_ERR_ARG, "window %d: empty interval", w);
        for (int i = 0; i < nInd; ++i) {
            const int64_t wi = (int64_t)w * nInd + i;
            const int64_t b0 = hb->wi_slot_off[wi];
            const int n_good = hb->wi_n_good[wi];
            if (n_good > 0) {
                const int r_first = hb->slot_read[b0];
                const int64_t rlen = hb->read_seq_off[r_first + 1] - hb->read_seq_off[r_first];
                const int rate = (int)std::max<int64_t>(1, rlen * n_good / size / target_coverage);
                for (int t = 0; t < n_good; t += rate) slots.push_back(hb->slot_read[b0 + t]);
            }
            slot_off[(size_t)wi + 1] = (int64_t)slots.size();
        }
    }
    PlbWindowBatch g = *hb;
    g.n_slots = slot_off[(size_t)nwi];
    g.wi_slot_off = slot_off.data();
    g.slot_read = slots.data();
    g.wi_n_good = zero.data();
    g.wi_n_bad = zero.data();
    g.max_variants = 0;
    g.win_n_var = nullptr;
    g.hap_var_mask = nullptr;
    g.var_prior = nullptr;
    PlbDeviceBatch* db = nullptr;
    if ((rc = plb_batch_upload(c, &g, &db))) return rc;
    cudaStream_t st = c->stream;
    Block out{nullptr, 0};
    const bool modes = opt.calc_flank_score != 0;
    const size_t o_h1 = ((size_t)n_pairs * 8 + 255) & ~(size_t)255, o_h2 = o_h1 + (((size_t)n_pairs * 4 + 255) & ~(size_t)255);
    rc = block_get(c, o_h2 + (size_t)n_pairs * 4 + 256, &out);
    cudaError_t e = cudaSuccess;
    if (!rc) e = cudaMemcpyAsync((uint8_t*)out.p + o_h1, hap1, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st);
    if (!rc && e == cudaSuccess) e = cudaMemcpyAsync((uint8_t*)out.p + o_h2, hap2, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st);
    if (!rc && modes) rc = mode_queues(c, db, false);
    if (!rc) rc = launch_windows(c, db, db->chunks[0], &opt, nullptr, nullptr, st, false, modes ? db->mq : db->q);
    if (!rc && e == cudaSuccess) {
        k_pair_score<<<(n_pairs + 3) / 4, 128, 0, st>>>(db->d, db->ll_scratch, n_pairs, (const int32_t*)((uint8_t*)out.p + o_h1),
                                                        (const int32_t*)((uint8_t*)out.p + o_h2), (double*)out.p);
        rc = launch_check(c, "k_pair_score");
    }
    if (!rc && e == cudaSuccess) e = cudaMemcpyAsync(score_out, out.p, (size_t)n_pairs * 8, cudaMemcpyDeviceToHost, st);
    const cudaError_t e2 = cudaStreamSynchronize(st);
    if (out.p) block_put(c, out);
    plb_batch_free(c, db);
    if (rc) return rc;
    if (e != cudaSuccess || e2 != cudaSuccess)
        return set_err(PLB_ERR_CUDA, "plb_best_score_genotypes_host: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return PLB_OK;
}
