// plb_stage.cuh — read staging on the host: BAM records -> the engine's read pool.
//
// What the reference does per record between the BAM file and the window model
//   ReadIterator.get          src/cython/htslibWrapper.pyx:328-406   record -> cAlignedRead (ASCII bases, soft-clip adjusted pos)
//   addReadToBuffer           src/cython/cwindow.pyx:560-595         good / bad list
//   checkAndTrimRead          src/cython/cwindow.pyx:332-481         filters, QC-fail flag, quality trimming
//   ReadArray.setWindowPointers  src/cython/cwindow.pyx:208-236      the reads of a window
// is done here in one pass over arrays of record fields, and the bases go from BAM's 4-bit nibbles straight into the
// 2-bit packed pool the GPU path takes (PLB_SEQ_2BIT) - no ASCII copy of a read ever exists.  Host code only: this is
// staging, nothing here touches a likelihood.  Pinned against the reference's own bamReadBuffer on every record of its
// test BAM (tests/golden/n3_ref.npz, made by tests/golden/make_n3_fixture.py through oracle/_ref).
#pragma once

namespace {

enum { kLowQualBases = 0, kUnmappedRead, kMateUnmapped, kMateDistant, kSmallInsert, kDuplicate, kLowMapQual };   // cwindow.pyx:40-46

// SAM flag bits (htslibWrapper.pxd:233-296)
constexpr unsigned kFPaired = 0x1, kFProper = 0x2, kFUnmapped = 0x4, kFMateUnmapped = 0x8, kFReverse = 0x10,
                   kFMateReverse = 0x20, kFSecondary = 0x100, kFQcFail = 0x200, kFDuplicate = 0x400;

struct StagedPrev {   // the fields of theLastRead the duplicate rule looks at
    bool have = false;
    int pos = 0, rlen = 0, mate_pos = 0;
};

// checkAndTrimRead for one read.  q = its qualities (trimmed in place), L = rlen.  Returns 1 = good list.
int check_and_trim(unsigned& flag, int mapq, int chrom, int mate_chrom, int pos, int mate_pos, int isize, uint8_t* q, int L,
                   const uint32_t* cigar, int n_cigar, const StagedPrev& last, const PlbReadFilterOptions& o, int32_t* counts) {
    if (flag & kFSecondary) {
        flag |= kFQcFail;
        return 0;
    }
    if (mapq < o.min_map_qual) {
        counts[kLowMapQual]++;
        flag |= kFQcFail;
        return 0;
    }
    int n_low = 0;
    for (int i = 0; i < L; ++i) n_low += (int)(int8_t)q[i] < o.min_base_qual;   // the reference's qualities are signed chars
    if (L - n_low < o.min_good_qual_bases) {
        counts[kLowQualBases]++;
        flag |= kFQcFail;
        return 0;
    }
    if (flag & kFUnmapped) {
        counts[kUnmappedRead]++;
        flag |= kFQcFail;
        return 0;
    }
    const bool paired = flag & kFPaired;
    if (o.filter_mate_unmapped && paired && (flag & kFMateUnmapped)) {
        counts[kMateUnmapped]++;
        return 0;   // no QC-fail flag: such reads are still scored from the bad-read list
    }
    if (o.filter_mate_distant && paired && (chrom != mate_chrom || !(flag & kFProper))) {
        counts[kMateDistant]++;
        return 0;
    }
    const int abs_ins = isize < 0 ? -isize : isize;
    if (o.filter_small_insert && paired && isize != 0 && abs_ins < L) {
        counts[kSmallInsert]++;
        flag |= kFQcFail;
        return 0;
    }
    if (o.filter_duplicates) {
        bool dup = (flag & kFDuplicate) != 0;
        if (!dup && last.have && pos == last.pos && L == last.rlen) dup = paired ? last.mate_pos == mate_pos : true;
        if (dup) {
            counts[kDuplicate]++;
            flag |= kFQcFail;
            return 0;
        }
    }
    // usable: low-quality tail, overlapping mate, adapter read-through and soft clips get quality 0
    if (!(flag & kFReverse)) {
        for (int i = 1; i <= L; ++i) {
            if (i < o.trim_read_flank || (int)(int8_t)q[L - i] < 5) q[L - i] = 0;
            else break;
        }
    } else {
        for (int i = 0; i < L; ++i) {
            if (i < o.trim_read_flank || (int)(int8_t)q[i] < 5) q[i] = 0;
            else break;
        }
    }
    if (o.trim_overlapping == 1 && paired && abs_ins > 0 && !(flag & kFReverse) && (flag & kFMateReverse) && abs_ins < 2 * L) {
        const int64_t lim = std::min<int64_t>(L, (int64_t)2 * L - isize + 1);   // the signed insert size, as the reference has it
        for (int64_t i = 1; i <= lim; ++i) q[L - i] = 0;
    }
    if (o.trim_adapter == 1 && paired && abs_ins > 0 && abs_ins < L) {
        if (flag & kFReverse) {
            for (int i = 1; i < L - abs_ins + 1; ++i) q[L - i] = 0;
        } else {
            for (int i = abs_ins; i < L; ++i) q[i] = 0;
        }
    }
    if (o.trim_soft_clipped == 1) {
        int idx = 0;
        for (int k = 0; k < n_cigar; ++k) {
            const int op = (int)(cigar[k] & 0xF), len = (int)(cigar[k] >> 4);
            if (op == 0 || op == 1) {
                idx += len;
            } else if (op == 4) {
                for (int j = 0; j < len && idx < L; ++j) q[idx++] = 0;
            }
        }
    }
    return 1;
}

}  // namespace

extern "C" int plb_stage_reads_host(const PlbBamRecords* in, const PlbReadFilterOptions* opt, PlbStagedReads* out) {
    if (!in || !opt || !out) return set_err(PLB_ERR_ARG, "NULL argument");
    const int n = in->n;
    if (n < 0) return set_err(PLB_ERR_ARG, "bad record count");
    memset(out->counts, 0, sizeof out->counts);
    out->n_exc = 0;
    if (n == 0) return PLB_OK;
    if (!in->ref_id || !in->pos || !in->mapq || !in->flag || !in->mate_ref_id || !in->mate_pos || !in->tlen || !in->cigar_off ||
        !in->seq_off || !in->nib_off || !in->nib || !in->qual)
        return set_err(PLB_ERR_ARG, "NULL array in PlbBamRecords");
    if (!out->kept || !out->good || !out->read_pos || !out->read_end || !out->flag_out || !out->qual_out || !out->seq2)
        return set_err(PLB_ERR_ARG, "NULL array in PlbStagedReads");
    const int64_t nb = in->seq_off[n];
    memcpy(out->qual_out, in->qual, (size_t)nb);
    memset(out->seq2, 0, (size_t)((nb + 3) / 4 + 1));
    // the filters are sequential by nature (the duplicate rule looks at the previous read); packing is not
    StagedPrev last;
    for (int i = 0; i < n; ++i) {
        const int64_t b0 = in->seq_off[i];
        const int L = (int)(in->seq_off[i + 1] - b0);
        if (L < 0 || L > 32767) return set_err(PLB_ERR_SHAPE, "record %d: sequence length %d out of range", i, L);
        const int64_t c0 = in->cigar_off[i];
        const int nc = (int)(in->cigar_off[i + 1] - c0);
        if (nc < 0 || (nc > 0 && !in->cigar)) return set_err(PLB_ERR_ARG, "record %d: bad CIGAR offsets", i);
        unsigned flag = in->flag[i];
        out->flag_out[i] = (uint16_t)flag;
        out->kept[i] = out->good[i] = 0;
        out->read_pos[i] = out->read_end[i] = 0;
        // ReadIterator.get: records without sequence or without qualities are not reads (htslibWrapper.pyx:334-338)
        if (L == 0 || in->qual[b0] == 0xFF) continue;
        const uint32_t* cg = in->cigar + c0;
        // pos = first base of the READ: a leading soft clip is subtracted (.pyx:383-387); end = bam_endpos
        int start = in->pos[i];
        if (nc > 0 && (cg[0] & 0xF) == 4) start -= (int)(cg[0] >> 4);
        int64_t ref_len = 0;
        for (int k = 0; k < nc; ++k) {
            const int op = (int)(cg[k] & 0xF);
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_len += cg[k] >> 4;
        }
        const int end = in->pos[i] + (int)(ref_len > 0 ? ref_len : 1);
        const int ok = check_and_trim(flag, in->mapq[i], in->ref_id[i], in->mate_ref_id[i], start, in->mate_pos[i], in->tlen[i],
                                      out->qual_out + b0, L, cg, nc, last, *opt, out->counts);
        last.have = true;
        last.pos = start;
        last.rlen = L;
        last.mate_pos = in->mate_pos[i];
        out->kept[i] = 1;
        out->good[i] = (uint8_t)ok;
        out->read_pos[i] = start;
        out->read_end[i] = end;
        out->flag_out[i] = (uint16_t)flag;
    }
    // a switched-off filter reports -1, as the reference's filteredReadCountsByType does (cwindow.pyx:516-526)
    if (!opt->filter_duplicates) out->counts[kDuplicate] = -1;
    if (!opt->filter_mate_unmapped) out->counts[kMateUnmapped] = -1;
    if (!opt->filter_mate_distant) out->counts[kMateDistant] = -1;
    if (!opt->filter_small_insert) out->counts[kSmallInsert] = -1;
    // bases: BAM nibbles -> 2-bit codes + exceptions, records side by side in the pool at their base offsets
    static const char* const kNib = "=ACMGRSVTWYHKDBN";   // htslibWrapper.pyx:414-416
    std::vector<std::vector<std::pair<int64_t, uint8_t>>> exc((size_t)host_threads());
#pragma omp parallel num_threads(host_threads()) if (nb > (1 << 18))
    {
        const int nt = omp_get_num_threads(), t = omp_get_thread_num();
        // records are dealt in contiguous blocks; a block boundary may split a packed byte between two threads, so the
        // few boundary bases are OR-ed in atomically
        const int r0 = (int)((int64_t)n * t / nt), r1 = (int)((int64_t)n * (t + 1) / nt);
        for (int i = r0; i < r1; ++i) {
            if (!out->kept[i]) continue;
            const int64_t b0 = in->seq_off[i];
            const int L = (int)(in->seq_off[i + 1] - b0);
            const uint8_t* nb4 = in->nib + in->nib_off[i];
            for (int k = 0; k < L; ++k) {
                const int code4 = (nb4[k >> 1] >> (4 * (1 - (k & 1)))) & 15;
                int c2;
                switch (code4) {
                    case 1: c2 = 0; break;
                    case 2: c2 = 1; break;
                    case 4: c2 = 2; break;
                    case 8: c2 = 3; break;
                    default:
                        c2 = 0;
                        exc[(size_t)t].push_back({b0 + k, (uint8_t)kNib[code4]});
                }
                if (c2) {
                    const int64_t j = b0 + k;
                    const uint8_t bits = (uint8_t)(c2 << (2 * (j & 3)));
                    if (k < 4 || k >= L - 4) {
#pragma omp atomic
                        out->seq2[j >> 2] |= bits;
                    } else {
                        out->seq2[j >> 2] |= bits;
                    }
                }
            }
        }
    }
    int64_t k = 0;
    for (auto& v : exc)
        for (auto& e : v) {
            if (k >= out->exc_cap || !out->exc_pos || !out->exc_chr)
                return set_err(PLB_ERR_SHAPE, "more than %lld bases outside ACGT: exception arrays too small", (long long)out->exc_cap);
            out->exc_pos[k] = e.first;
            out->exc_chr[k] = e.second;
            ++k;
        }
    out->n_exc = k;
    return PLB_OK;
}

extern "C" int plb_window_slices_host(int32_t n_reads, const int32_t* read_pos, const int32_t* read_end, int32_t n_windows,
                                      const int32_t* win_start, const int32_t* win_end, int32_t* lo_out, int32_t* hi_out) {
    if (n_reads < 0 || n_windows < 0 || (n_windows > 0 && (!win_start || !win_end || !lo_out || !hi_out)) ||
        (n_reads > 0 && (!read_pos || !read_end)))
        return set_err(PLB_ERR_ARG, "NULL / bad argument");
    int longest = 0;   // ReadArray.__longestRead: the largest end - pos appended so far (cwindow.pyx:169-174)
    for (int i = 0; i < n_reads; ++i) longest = std::max(longest, read_end[i] - read_pos[i]);
    // bisectReadsLeft (cwindow.pyx:276-300), step for step: file order is BAM-position order, while pos has a leading soft
    // clip subtracted, so a list can be slightly out of order - the reference bisects it anyway and so do we
    auto bisect = [&](int test) {
        int low = 0, high = n_reads;
        while (low < high) {
            const int mid = (low + high) / 2;
            if (read_pos[mid] < test) low = mid + 1;
            else high = mid;
        }
        return low;
    };
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (n_windows > 4096)
    for (int w = 0; w < n_windows; ++w) {
        if (n_reads == 0) {
            lo_out[w] = hi_out[w] = 0;
            continue;
        }
        const int start = win_start[w], end = win_end[w];
        const int first = std::max(1, start - longest);
        int lo = bisect(first);
        const int hi = bisect(end);
        while (lo < n_reads && read_end[lo] <= start) ++lo;
        lo_out[w] = lo;
        hi_out[w] = std::max(lo, std::min(hi, n_reads));
    }
    return PLB_OK;
}
