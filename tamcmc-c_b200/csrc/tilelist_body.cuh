// tilelist_body.cuh -- one warp builds the component lists of one (star, chain, tile) work item.
// Included by tilelist.cu (standalone kernel) and whittle.cu (builder warps of the fused kernel).
#pragma once
#include "tamcmc_dev.h"
#include "kernels.h"

namespace tamcmc_tl {

__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ int warp_excl_scan(int v, int lane)
{
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += o; }
    return x - v;
}

// classifies the chain's modes against the tile (bit-exact windows from the expander) and writes
//   FastEntry fast[TF] | ModeHdr hdr[TH] | GenEntry gen[TG] | SegDesc seg[..]   into the list pool,
// and the list descriptor into the tile record
__device__ void build_tile_lists(const TileListArgs& A, unsigned int item, int lane)
{
    const int sc = (int)(item / (unsigned)A.tiles_stride);
    const int tile = (int)(item - (unsigned)sc * (unsigned)A.tiles_stride);
    const StarDesc* sd = A.stars + sc / A.Nchains;
    TileRec* tr = A.tilerec + item;
    const int nmodes = sd->nmodes_cap;
    const int lb0 = tile * TAMCMC_TILE;
    const int g0 = sd->bin0 + lb0;
    const int gend = g0 + min(TAMCMC_TILE, sd->Nloc - lb0);
    const double xc = tr->xc;
    const bool asym = A.asym_flag[sc] != 0;
    const ModeRec* modes = A.modes + (size_t)sc * A.modes_stride;
    const CompRec* comps = A.comps + (size_t)sc * A.modes_stride * TAMCMC_MAX_COMP_PER_MODE;

    // ---- sweep 1: totals ----
    int TF = 0, TG = 0, TH = 0;
    for (int base = 0; base < nmodes; base += 32) {
        const int mi = base + lane;
        int nf = 0, ng = 0;
        if (mi < nmodes) {
            const int4 h = *reinterpret_cast<const int4*>(modes + mi);     // {i0, i1, ncomp, nfast}
            if (h.z > 0 && h.x < gend && h.y > g0) {
                nf = (h.x <= g0 && h.y >= gend) ? (h.w & 0xffff) : 0;
                ng = h.z - nf;
            }
        }
        TF += nf; TG += ng; TH += (asym && nf > 0) ? 1 : 0;
    }
    TF = warp_sum(TF); TG = warp_sum(TG); TH = warp_sum(TH);
    const int nseg_cap = 1 + TF / (TAMCMC_CAPF - 7 * 32 > 0 ? TAMCMC_CAPF - 7 * 32 : 1) + TG / 3 + TH / (TAMCMC_CAPH - 32 > 0 ? TAMCMC_CAPH - 32 : 1);
    const unsigned long long bytes = 32ull * TF + 32ull * TH + 64ull * TG + 32ull * nseg_cap;
    unsigned long long off = 0;
    if (lane == 0) off = atomicAdd(&A.qctl->pool_cursor, (bytes + 127ull) & ~127ull);
    off = __shfl_sync(0xffffffffu, off, 0);
    if (off + bytes > A.pool_bytes) {
        if (lane == 0) { atomicExch(&A.qctl->overflow, 1u); tr->nseg = 0; tr->TF = tr->TG = tr->TH = 0; tr->s0_nf = tr->s0_nh = tr->s0_ng = 0; tr->s0_wide = 0; tr->pool_off = 0; }
        return;
    }
    FastEntry* fast = reinterpret_cast<FastEntry*>(A.pool + off);
    ModeHdr* hdr = reinterpret_cast<ModeHdr*>(A.pool + off + 32ull * TF);
    GenEntry* gen = reinterpret_cast<GenEntry*>(A.pool + off + 32ull * TF + 32ull * TH);
    SegDesc* segs = reinterpret_cast<SegDesc*>(A.pool + off + 32ull * TF + 32ull * TH + 64ull * TG);

    // ---- sweep 2: emit.  (cf, cg, ch) = entries written so far; (f0, g0s, h0) = start of the open segment ----
    int cf = 0, cg = 0, ch = 0, f0 = 0, g0s = 0, h0 = 0, nseg = 0;
    int s0nf = 0, s0nh = 0, s0ng = 0, s0wide = 0, seg_wide = 0;
    for (int base = 0; base < nmodes; base += 32) {
        const int mi = base + lane;
        int ncomp = 0, nfast = 0, ngen = 0, i0 = 0, i1 = 0, nfast_rec = 0, mwide = 0, wbit = 0;
        if (mi < nmodes) {
            const int4 h = *reinterpret_cast<const int4*>(modes + mi);
            if (h.z > 0 && h.x < gend && h.y > g0) {
                ncomp = h.z; nfast_rec = h.w & 0xffff; i0 = h.x; i1 = h.y;
                nfast = (h.x <= g0 && h.y >= gend) ? nfast_rec : 0;
                wbit = (h.w >> 16) & 1;
                mwide = (nfast > 0) ? wbit : 0;
                ngen = ncomp - nfast;
            }
        }
        const int hh = (asym && nfast > 0) ? 1 : 0;
        int sub_lo = 0;
        while (sub_lo < 32) {
            int sub_hi = 32;
            bool mine = lane >= sub_lo;
            int tf = warp_sum(mine ? nfast : 0), tg = warp_sum(mine ? ngen : 0), th = warp_sum(mine ? hh : 0);
            if (tg > TAMCMC_CAPG) {
                // too many general entries for one segment: 3 modes at a time (3 x 7 <= CAPG)
                sub_hi = sub_lo + 3;
                mine = lane >= sub_lo && lane < sub_hi;
                tf = warp_sum(mine ? nfast : 0); tg = warp_sum(mine ? ngen : 0); th = warp_sum(mine ? hh : 0);
            }
            if ((cf - f0) + tf > TAMCMC_CAPF || (cg - g0s) + tg > TAMCMC_CAPG || (ch - h0) + th > TAMCMC_CAPH) {
                // close the open segment
                if (lane == 0) { SegDesc d; d.f0 = f0; d.nf = cf - f0; d.h0 = h0; d.nh = ch - h0; d.g0 = g0s; d.ng = cg - g0s; d.wide = seg_wide; d.pad1 = 0; segs[nseg] = d; }
                if (nseg == 0) { s0nf = cf - f0; s0nh = ch - h0; s0ng = cg - g0s; s0wide = seg_wide; }
                nseg++; f0 = cf; g0s = cg; h0 = ch; seg_wide = 0;
            }
            seg_wide |= __any_sync(0xffffffffu, mine && mwide) ? 1 : 0;
            const int mf = mine ? nfast : 0, mg = mine ? ngen : 0, mh = mine ? hh : 0;
            const int of = cf + warp_excl_scan(mf, lane), og = cg + warp_excl_scan(mg, lane), oh = ch + warp_excl_scan(mh, lane);
            if (mine && ncomp > 0) {
                const CompRec* cp = comps + (size_t)mi * TAMCMC_MAX_COMP_PER_MODE;
                double qa = 0.0, qb = 1.0, qc = 0.0;
                if (asym || ngen > 0) { const ModeRec* mr = modes + mi; qa = mr->qa; qb = mr->qb0 + xc * qa; qc = mr->qc; }
                if (hh) { ModeHdr m; m.qa = qa; m.qb = qb; m.qc = qc; m.begin = of - f0; m.count = nfast; hdr[oh] = m; }
#pragma unroll
                for (int k0 = 0; k0 < 8; k0 += 4) {
                    double cnu[4], cs[4], ca[4];
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) { const int k = k0 + kk; if (k < ncomp) { cnu[kk] = cp[k].nu; cs[kk] = cp[k].s; ca[kk] = cp[k].a; } }
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        const int k = k0 + kk;
                        if (k < ncomp) {
                            const double cc = -(cnu[kk] - xc) * cs[kk];
                            if (k < nfast) { FastEntry fe; fe.s = cs[kk]; fe.c = cc; fe.a = ca[kk]; fe.pad = 0.0; fast[of + k] = fe; }
                            else {
                                // components are stored FAST-first; a FAST one lands here only on a window edge
                                const bool ff = k < nfast_rec;
                                GenEntry ge;
                                ge.s = cs[kk]; ge.c = cc; ge.aadd = ff ? ca[kk] : 1.0; ge.num = ff ? 1.0 : ca[kk];
                                // window in tile-local bins, clamped to the tile; bit 30 of hi: the entry needs an exponent
                                // renormalisation after every merge (general form, or a WIDE-range mode)
                                ge.qa = qa; ge.qb = qb; ge.qc = qc; ge.lo = max(i0 - g0, 0);
                                ge.hi = min(i1 - g0, TAMCMC_TILE) | ((!ff || wbit) ? (1 << 30) : 0);
                                gen[og + (k - nfast)] = ge;
                            }
                        }
                    }
                }
            }
            cf += tf; cg += tg; ch += th;
            sub_lo = sub_hi;
        }
    }
    // last segment (possibly empty: a tile no mode touches still has its background and Whittle terms)
    if (lane == 0) { SegDesc d; d.f0 = f0; d.nf = cf - f0; d.h0 = h0; d.nh = ch - h0; d.g0 = g0s; d.ng = cg - g0s; d.wide = seg_wide; d.pad1 = 0; segs[nseg] = d; }
    if (nseg == 0) { s0nf = cf - f0; s0nh = ch - h0; s0ng = cg - g0s; s0wide = seg_wide; }
    nseg++;
    if (lane == 0) {
        tr->pool_off = off; tr->nseg = nseg; tr->TF = TF; tr->TH = TH; tr->TG = TG;
        tr->s0_nf = s0nf; tr->s0_nh = s0nh; tr->s0_ng = s0ng; tr->s0_wide = s0wide;
    }
}

}  // namespace tamcmc_tl
