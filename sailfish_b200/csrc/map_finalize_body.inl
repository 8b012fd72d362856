// map_finalize_body.inl -- the body of k_finalize_reads, included once per kernel entry point (map.cu) with SFB_FIN_BIAS 0 or 1.
// Plain textual inclusion (no device-function wrapper).  SFB_FIN_BIAS 1 adds what --biasCorrect / --gcBiasCorrect collect per hit
// (SailfishQuantify.cpp:255-287, :372-389, :555-583): the read-start context of the read's first hit that has one (p.bias_val,
// sampled in read order by k_bias_select) and the GC percentage of every properly paired hit that lies inside its transcript
// (s_gc, the CTA's 101-bin histogram, declared by the including kernel).
//
// A lane owns one read per round; its hit lists live in shared memory (Scratch, map.cu).  The class upsert at the end of a round is
// warp-aggregated: lanes whose labels are equal (same XXH64, verified member by member) elect a leader that adds the group's count
// with ONE table probe and ONE atomic (EquivalenceClassBuilder::addGroup, include/EquivalenceClassBuilder.hpp:90-108, called once
// per read by the reference).
    const uint64_t gtid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t cap = p.cap;
    const Scratch scr{(threadIdx.x >> 5) * (N_REGIONS * FIN_S * 32) + lane, p.scratch + gtid, p.n_threads_total, cap + 1};
    const bool paired = p.n_mates == 2;
    const int ns = 2 * p.n_mates;
    uint32_t c_obs = 0, c_map = 0, c_hits = 0, c_ub = 0, c_fw = 0, c_rc = 0;     // per thread and chunk (<= 2^22 fragments x 200 hits / 10^5 threads)
    Interval ivs[4][MAX_IV];
    int niv[4];
    uint64_t score[4];

    // retry pass (p.retry_in != nullptr): the class table or the label arena was full when these reads were finalized; the table has
    // grown since (sfb_eq_grow, map.cu) and only their class upsert is repeated -- they were counted the first time.
    // heavy pass (p.heavy_in != nullptr): the reads the main pass set aside because one of their seeds has a bucket of more than
    // SFB_HEAVY_CNT positions (repeats, paralog families).  Walking such a bucket is hundreds of dependent loads by ONE lane; in the
    // main pass the other 31 lanes of the warp waited for it (ncu on the paralog set: 2 of 32 lanes active, 7.2 ms per 2 M reads
    // against 0.6 ms without such reads).  Set aside and finalized together, every lane of a warp has such a read.
    const bool retry = p.retry_in != nullptr;
    const bool heavy_pass = p.heavy_in != nullptr;
    // the heavy list is kept in three bins by the size of the read's largest bucket (p.n_reads entries apart, counted in
    // p.next_read[2..4]) and walked from the largest bin down, so that the reads of a warp cost about the same
    const uint64_t hv2 = heavy_pass ? (uint64_t)__ldcg(p.next_read + 4) : 0, hv1 = heavy_pass ? (uint64_t)__ldcg(p.next_read + 3) : 0;
    const uint64_t hv0 = heavy_pass ? (uint64_t)__ldcg(p.next_read + 2) : 0;
    const uint64_t n_work = retry ? p.n_retry : heavy_pass ? hv0 + hv1 + hv2 : p.n_reads;
    for (;;) {
        unsigned long long base_idx = 0;
        if (lane == 0) base_idx = atomicAdd(p.next_read + 1, 32ULL);
        base_idx = __shfl_sync(0xffffffffu, base_idx, 0);
        if (base_idx >= n_work) break;
        bool have_read = base_idx + lane < n_work;
        uint64_t ri = 0;
        if (have_read) {
            const uint64_t w = base_idx + lane;
            if (retry) ri = p.retry_in[w];
            else if (heavy_pass) ri = w < hv2 ? p.heavy_in[2 * p.n_reads + w] : w < hv2 + hv1 ? p.heavy_in[p.n_reads + (w - hv2)] : p.heavy_in[w - hv2 - hv1];
            else ri = w;
        }
        bool mapped = false;
        uint32_t lab_n = 0;
        uint32_t len1 = 0, len2 = 0;
        ReadG rg[2];
        if (have_read) {
            for (int mt = 0; mt < p.n_mates; ++mt) {
                const uint64_t fm = ri * p.n_mates + mt;
                const uint32_t me = p.meta[fm];
                rg[mt].pk = p.pk + fm * p.rwp; rg[mt].pkn = p.pkn + fm * p.rwp; rg[mt].len = me & 0xFFFFu; rg[mt].has_n = (me >> 16) & 1u;
            }
            len1 = rg[0].len;
            len2 = paired ? rg[1].len : 0;
            // the interval counts of all scans in one load, and the first interval of every scan loaded before the counts are known
            // (the hand-over arrays are allocated for MAX_IV intervals per scan, so the load is always in bounds)
            const uint32_t nv = paired ? *reinterpret_cast<const uint32_t*>(p.niv + ri * 4) : *reinterpret_cast<const uint16_t*>(p.niv + ri * 2);
            unsigned long long iv0[4]; uint32_t mk0[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (q < ns) { iv0[q] = p.iv[(ri * ns + q) * MAX_IV]; mk0[q] = p.ivmask[(ri * ns + q) * MAX_IV]; }
            }
            uint32_t big = 0;                                   // largest bucket among this read's seeds
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (q >= ns) continue;
                niv[q] = (nv >> (8 * q)) & 0xFFu;
                uint64_t sc = 0;
                if (niv[q] > 0) { ivs[q][0] = unpack_iv(iv0[q]); ivs[q][0].mask = mk0[q]; sc = ivs[q][0].m; big = ivs[q][0].cnt > big ? ivs[q][0].cnt : big; }
                for (int e = 1; e < niv[q]; ++e) {
                    ivs[q][e] = unpack_iv(p.iv[(ri * ns + q) * MAX_IV + e]);
                    ivs[q][e].mask = p.ivmask[(ri * ns + q) * MAX_IV + e];
                    sc += ivs[q][e].m;
                    big = ivs[q][e].cnt > big ? ivs[q][e].cnt : big;
                }
                score[q] = sc;
            }
            if (p.heavy_out && big > SFB_HEAVY_CNT) {          // main pass: not now
                const uint32_t bin = big <= 64u ? 0u : big <= 256u ? 1u : 2u;
                p.heavy_out[bin * p.n_reads + atomicAdd(p.next_read + 2 + bin, 1ULL)] = (uint32_t)ri;
                have_read = false;
            }
        }
        if (have_read) {
            uint32_t nL = 0, nR = 0, wL = R_LEFT, wR = R_RIGHT;
            bool okL, okR = true;
            okL = collect(p.ix, rg[0], paired, cap, ivs[0], niv[0], score[0], ivs[1], niv[1], score[1], scr, R_LEFT, nL, wL, p.pool);   // paired: strict check (:192-202)
            if (paired) {
                if (okL && wL != R_LEFT) {            // the projection regions are about to be reused by the right mate
                    for (uint32_t i = 0; i < nL; ++i) scr.set(R_LEFT, i, scr.get(wL, i));
                    wL = R_LEFT;
                }
                okR = collect(p.ix, rg[1], true, cap, ivs[2], niv[2], score[2], ivs[3], niv[3], score[3], scr, R_RIGHT, nR, wR, p.pool);
            }
            const bool overflow = !okL || !okR;
            LabelAcc acc(scr, p.enforce_compat != 0);
            uint32_t n_joint = 0;
            int32_t fl = -1;
#if SFB_FIN_BIAS
            int32_t bsample = -1;
#endif
            if (!paired) {
                // SailfishQuantify.cpp:530-631
                n_joint = overflow ? 0 : nL;
                c_ub += (overflow || n_joint > 0) ? 1 : 0;
                for (uint32_t i = 0; i < n_joint; ++i) {
                    const unsigned long long h = scr.get(wL, i);
#if SFB_FIN_BIAS
                    if (p.bias_seq && bsample < 0) bsample = hit_read_start_index(p.ix, hit_tid(h), hit_pos(h), hit_fwd(h), len1);
#endif
                    const bool compat = p.ignore_compat ? true : compat_single(p.lib_fmt, hit_fwd(h), 0);
                    acc.add(hit_tid(h), compat, hit_fwd(h));
                }
            } else if (!overflow) {
                // mergeLeftRightHits[Fuzzy] (call sites :204-213): one joint hit per transcript present in both lists
                uint32_t i = 0, j = 0, n_pairs = 0;
                while (i < nL && j < nR) {
                    const uint32_t tl = hit_tid(scr.get(wL, i)), tr = hit_tid(scr.get(wR, j));
                    if (tl < tr) ++i; else if (tr < tl) ++j;
                    else { ++n_pairs; ++i; while (i < nL && hit_tid(scr.get(wL, i)) == tl) ++i; while (j < nR && hit_tid(scr.get(wR, j)) == tl) ++j; }
                }
                if (n_pairs > 0) {
                    n_joint = n_pairs;
                    c_ub += 1;
                    i = 0; j = 0;
                    while (i < nL && j < nR) {                                              // :341-369
                        const unsigned long long hl = scr.get(wL, i), hr = scr.get(wR, j);
                        const uint32_t tl = hit_tid(hl), tr = hit_tid(hr);
                        if (tl < tr) { ++i; continue; }
                        if (tr < tl) { ++j; continue; }
                        const int32_t pl = hit_pos(hl), pr = hit_pos(hr);
                        const bool fl_ = hit_fwd(hl), fr_ = hit_fwd(hr);
#if SFB_FIN_BIAS
                        if (p.bias_seq && bsample < 0) bsample = hit_read_start_index(p.ix, tl, pl, fl_, len1);
                        if (p.bias_gc && !retry) {                                      // :375-388
                            const int32_t start = pl < pr ? pl : pr;
                            const int32_t e1 = pl + (int32_t)len1, e2 = pr + (int32_t)len2;
                            const int32_t stop = e1 > e2 ? e1 : e2;                      // start + fragLen
                            const uint64_t t0 = p.ix.txp_start[tl];
                            if (start > 0 && stop < (int32_t)(p.ix.txp_end[tl] - t0)) atomicAdd(s_gc + b_gc_frac_range(p.ix.words, t0, start, stop), 1u);
                        }
#endif
                        bool compat = p.ignore_compat != 0;
                        if (!compat) {
                            const uint32_t e1 = fl_ ? (uint32_t)pl : (uint32_t)pl + len1;
                            const uint32_t e2 = fr_ ? (uint32_t)pr : (uint32_t)pr + len2;
                            compat = compat_paired(p.lib_fmt, (int32_t)e1, fl_, len1, (int32_t)e2, fr_, len2, p.allow_dovetail != 0);
                        }
                        acc.add(tl, compat, fl_);
                        if (n_pairs == 1) {
                            const int32_t fs = pl < pr ? pl : pr;
                            const int32_t e1 = pl + (int32_t)len1, e2 = pr + (int32_t)len2;
                            fl = (e1 > e2 ? e1 : e2) - fs;
                        }
                        ++i; while (i < nL && hit_tid(scr.get(wL, i)) == tl) ++i; while (j < nR && hit_tid(scr.get(wR, j)) == tl) ++j;
                    }
                } else if (!p.strict_intersect && nL + nR > 0) {
                    // orphans: left block then right block, merged by transcript id (:231-246), left first on ties
                    n_joint = nL + nR;
                    c_ub += 1;
                    if (n_joint > cap) n_joint = 0;                                        // :217
                    else if (!p.allow_orphans) { /* :226 joint hits discarded, n_joint keeps counting them below */ }
                    if (n_joint > 0 && p.allow_orphans) {
                        i = 0; j = 0;
                        while (i < nL || j < nR) {                                          // :289-340
                            bool takeL;
                            if (i >= nL) takeL = false; else if (j >= nR) takeL = true;
                            else takeL = hit_tid(scr.get(wL, i)) <= hit_tid(scr.get(wR, j));
                            const unsigned long long h = takeL ? scr.get(wL, i++) : scr.get(wR, j++);
                            const int ms = takeL ? 1 : 2;
                            const bool fwd = hit_fwd(h);
#if SFB_FIN_BIAS
                            if (p.bias_seq && bsample < 0) bsample = hit_read_start_index(p.ix, hit_tid(h), hit_pos(h), fwd, ms == 1 ? len1 : len2);
#endif
                            const bool compat = p.ignore_compat ? true : compat_single(p.lib_fmt, fwd, ms);
                            const bool fwdHit = takeL ? fwd : !fwd;
                            acc.add(hit_tid(h), compat, fwdHit);
                        }
                    } else if (n_joint > 0) {
                        n_joint = 0;                                                        // :226 jointHits.clear()
                    }
                }
            } else {
                c_ub += 1;                                                                  // an overflowed mate did have hits
            }
            if (acc.n > 0 && (acc.haveCompat || !p.enforce_compat)) {
                mapped = true;
                lab_n = acc.n;
                c_fw += acc.fw; c_rc += acc.rc;
            }
            if (p.fld_val && !retry) {
                const bool elig = paired && n_joint == 1 && fl >= 0 && mapped && (uint32_t)fl < p.max_frag_len;   // :419-434
                p.fld_val[ri] = elig ? (int16_t)fl : (int16_t)-1;
            }
#if SFB_FIN_BIAS
            if (p.bias_val && !retry) p.bias_val[ri] = (int16_t)bsample;
#endif
            c_obs += 1; c_map += mapped ? 1 : 0; c_hits += n_joint;
        }
        bool failed = false;                               // the upsert this lane's read took part in found the table or the arena full
        // ---- warp-aggregated class upsert (the warp is convergent here) ----
        const unsigned map_m = __ballot_sync(0xffffffffu, mapped);
        if (mapped) {
            auto get = [&](uint32_t j) { return (uint32_t)scr.get(R_LABEL, j); };
            const uint64_t h = xxh64_words(get, lab_n, 0);                                 // TranscriptGroup.cpp:9-12
            const unsigned grp = __match_any_sync(map_m, h);
            const int leader = __ffs(grp) - 1;
            const uint32_t lead_n = __shfl_sync(grp, lab_n, leader);
            bool same = lead_n == lab_n;
            if (same && (int)lane != leader) {                                               // equal hashes are not yet equal labels
                const int d = leader - (int)lane;
                for (uint32_t j = 0; j < lab_n && same; ++j) same = (uint32_t)scr.peer(d, R_LABEL, j) == get(j);
            }
            const unsigned agree = __ballot_sync(map_m, same) & grp;
            bool ok = true;
            if ((int)lane == leader) ok = eq_upsert(p.tb, lab_n, get, (unsigned long long)__popc(agree), h);
            else if (!same) ok = eq_upsert(p.tb, lab_n, get, 1ULL, h);
            const bool lead_ok = __shfl_sync(grp, (int)ok, leader) != 0;   // the members of a group share their leader's fate
            failed = same ? !lead_ok : !ok;
            if (failed && p.retry_out) p.retry_out[atomicAdd(p.tb.cursor + 4, 1ULL)] = (uint32_t)ri;
        }
        __syncwarp();                                      // the next round overwrites the label region other lanes may still be comparing
    }
    // warp-reduce the six counters (ReadExperiment.hpp:74-97), one atomic per warp and counter
    const unsigned long long v[6] = {c_obs, c_map, c_hits, c_ub, c_fw, c_rc};
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        unsigned long long x = v[q];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) x += __shfl_xor_sync(0xffffffffu, x, m);
        if (lane == 0 && x && !retry) atomicAdd(p.counters + q, x);
    }
