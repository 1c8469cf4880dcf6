// The chain stage for three and more channels: the same wavefront over consecutive blocks of a stream as chain_kernel
// (chain.cuh), with the CHANNELS of a block spread over lanes.  With one lane per block a step costs the sum of all
// channels' work -- an 8-channel step ran 3.2 us, and because block m+1 can only follow block m `lag` steps behind, a single
// long stream (BASELINE configs[4]: 1 h, 96 kHz, 8 channels, 400 blocks/s) takes blocks x lag x step time whatever the GPU
// has to offer.  The follower channels of a bin depend only on the maximum-energy channel's new output at the same bin (W#48
// 9458-9873), and the preliminary prediction (S5) is per channel: so a block gets CPL = 4 or 8 lanes, lane (block, channel)
//   - runs S5 of its own channel (previous block's output of that channel: from CPL lanes below by shuffle),
//   - takes the four neighbour values of the chain from the lane of the maximum channel by shuffle and computes the
//     phase sum and the maximum channel's output redundantly (same operations on the same values in every lane),
//   - derives its own channel's output from it.
// SPLIT (few streams: a lone stream's time is blocks x lag x the time of ONE step): S5 at bin q and S6 at bin k of a step
// both depend only on the step before -- S6(t) reads the S5 value of bin k + longStep = q - 1, written in step t - 1 -- so
// they run side by side on twice the warps: warps [0, nW) do S6, warps [nW, 2 nW) S5 of the same (block, channel) lanes,
// exchanging through the shared-memory rings under the barrier every step already has.
// Why the twins never touch the same ring slot in the same step (the barrier at the top of a step is the only ordering between
// them):  in step t a block is at S6 bin k and S5 bin q = k + longStep + 1.
//   ringO (outputs, written by S6 at slot k): the S5 twin of the NEXT block reads its predecessor's slot q' = k - 1 of that
//     column -- last step's write; this step's write goes to slot k, and slot k - 1 is not rewritten before step t + RO - 1.
//   ringN (S5 values, written by S5 at slot q): S6 reads slots k + 1 = q - longStep and k + longStep = q - 1 -- older writes, the
//     younger of them from step t - 1; RN > longStep keeps q, q - 1 and q - longStep apart.
//   record stage of step t - 1: read by both twins before the barrier of step t, refilled by the copy issued after it.
// A warp holds 32 / CPL blocks, a CTA of 8 warps 32 or 64: one or two record groups.  The rows a warp needs in a step
// (its blocks' rows of one diagonal) are contiguous and arrive by cp.async.bulk into a ring of kWideStages stages; the
// row pitch is 8 (mod 32) floats so that the lanes' field reads of different blocks fall into different banks.
#pragma once
#include "chain.cuh"

namespace bs {

#ifndef BS_WIDE_STAGES
#define BS_WIDE_STAGES 4
#endif
constexpr int kWideStages = BS_WIDE_STAGES;
#ifndef BS_WIDE_TILE
#define BS_WIDE_TILE 8
#endif
#ifndef BS_WIDE_AHEAD
#define BS_WIDE_AHEAD 2
#endif
constexpr int kWideTile = BS_WIDE_TILE;    // bins per relay tile.  A long single stream is a chain of hundreds of hand-overs from CTA to
constexpr int kWideAhead = BS_WIDE_AHEAD;  // CTA, and a follower runs kWideTile + kWideAhead bins behind its predecessor where blocks inside
                                           // a CTA are lag = longStep + 2 apart: a tile is asked for kWideAhead steps before its first bin is
                                           // due (time for the poll + the copy), and the predecessor has to have finished it by then
static_assert(kWideTile >= 4 && (kWideTile & (kWideTile - 1)) == 0 && kWideAhead >= 1 && kWideAhead < kWideTile, "relay tile");
BS_HHD constexpr int wide_cpl(int C) { return C <= 4 ? 4 : 8; }                  // lanes per block
BS_HHD constexpr int wide_bpw(int C) { return 32 / wide_cpl(C); }                // blocks per warp
BS_HHD constexpr int wide_pass_blocks(int C, int warps) { return warps * wide_bpw(C); }   // blocks per CTA
BS_HHD size_t wide_smem_bytes(int C, int longStep, int warps) {
  const size_t perWarp = (size_t)kWideStages * wide_bpw(C) * nr_pitch(C) * sizeof(float) + (size_t)(chain_ring_n(longStep) + chain_ring_o(longStep)) * 32 * sizeof(cf);
  return warps * perWarp + 2 * (size_t)kWideTile * C * sizeof(cf) + 2 * (size_t)warps * C * sizeof(cf) + (size_t)warps * kWideStages * sizeof(unsigned long long) + 16;
}
// few streams (2 or 4 warps of blocks per CTA): S5 and S6 on separate warps, twice the threads
BS_HHD constexpr bool wide_split(int warps) { return warps <= 4; }
BS_HHD constexpr int wide_threads(int warps) { return (wide_split(warps) ? 64 : 32) * warps; }
// (warps of blocks per CTA: wide_warps_for, kernels.cuh)

// the phase sum of S6 at bin k (chain_bin / chain_fast, kernels.cuh / chain.cuh): no division in here, one form for both
__device__ __forceinline__ void chain_phase(const float *ra, int k, int B, int ls, cf oPrev, cf oLong, cf n1, cf nL, float &phRe, float &phIm) {
  phIm = (ra[1] * oPrev.re) + (ra[0] * oPrev.im); phRe = (ra[0] * oPrev.re) - (ra[1] * oPrev.im);
  if (!(k > 0)) { phIm = 0.f; phRe = 0.f; }
  {
    const float aIm = ((ra[2] * oLong.im) + phIm) + (ra[3] * oLong.re), aRe = ((ra[2] * oLong.re) + phRe) - (oLong.im * ra[3]);
    if (k >= ls) { phIm = aIm; phRe = aRe; }
  }
  {
    const float t4 = ra[4] * n1.re, t5 = ra[5] * n1.im, t8 = (ra[4] * n1.im) - (ra[5] * n1.re);
    const float aIm = t8 + phIm, aRe = (t4 + phRe) + t5;
    if (k < B - 1) { phIm = aIm; phRe = aRe; }
  }
  {
    const float t6 = ra[6] * nL.re, t7 = ra[7] * nL.im, t9 = ra[6] * nL.im, t10 = nL.re * ra[7];
    const float aIm = (t9 + phIm) - t10, aRe = (t6 + phRe) + t7;
    if (k < B - ls) { phIm = aIm; phRe = aRe; }
  }
}

template <int C, bool SPLIT>
__global__ void __launch_bounds__(32 * kChainWarps, 2) chain_wide_kernel(DevGeom g, DevTables T, const StreamDev *streams, const BlockRec *blocks,
                                                                     const BlockRec2 *blocks2, long long slot0, int nSlots, const cf *specIn,
                                                                     cf *specOut, StateDev st, int ctas, int *prog, int *err) {
  extern __shared__ float4 sm4[];
  constexpr int CPL = wide_cpl(C), BPW = wide_bpw(C), NRP = nr_pitch(C), SO = 9 + 5 * C, TL = kWideTile, NSTG = kWideStages;
  constexpr unsigned full = 0xffffffffu;
  const int nW = SPLIT ? blockDim.x >> 6 : blockDim.x >> 5, perPass = nW * BPW, nThreads = blockDim.x;
  __shared__ int ticket;
  if (ctas > 1) {   // relayed launch: logical CTA index = order of arrival (chain.cuh)
    if (threadIdx.x == 0) ticket = atomicAdd(prog, 1);
    __syncthreads();
  }
  const int bid = ctas > 1 ? ticket : (int)blockIdx.x;
  const int s = bid / ctas, cta = bid - s * ctas, lane = threadIdx.x & 31, tid = threadIdx.x;
  const bool roleS5 = SPLIT && (int)(threadIdx.x >> 5) >= nW, roleS6 = !roleS5;   // (without SPLIT every warp does both)
  const int warp = (int)(threadIdx.x >> 5) - (roleS5 ? nW : 0);
  const int bl = lane / CPL, cRaw = lane % CPL, c = cRaw < C ? cRaw : C - 1;   // idle lanes (C < CPL) shadow the last channel, store nothing
  const bool chOK = cRaw < C;
  const int jb = warp * BPW + bl;                            // this lane's block within the CTA's pass
  const StreamDev sd = streams[s];
  const int B = g.B, ls = g.longStep, D = ls + 2, OA = 1, RN = chain_ring_n(ls), RO = chain_ring_o(ls), RMN = RN - 1, RMO = RO - 1;
  const int rows = rec_rows(B, ls);
  // shared memory: record stages [nW][NSTG][BPW rows], per warp ringN [RN][32] / ringO [RO][32] (one channel per lane), the
  // carried-state tiles, the hand-off slots between warps, the stage barriers
  float *stageAll = (float *)sm4;
  cf *rings = (cf *)(stageAll + (size_t)nW * NSTG * BPW * NRP);
  cf *ringN = rings + (size_t)warp * (RN + RO) * 32 + lane, *ringO = ringN + (size_t)RN * 32;
  cf *tile = rings + (size_t)nW * (RN + RO) * 32;          // [2][C][TL]
  cf *hand = tile + 2 * (size_t)C * TL;                    // [2][nW][C]
  unsigned long long *bars = (unsigned long long *)(hand + 2 * (size_t)nW * C) + NSTG * warp;
  long long nv = sd.nBlocks - slot0; if (nv > nSlots) nv = nSlots;
  if (nv <= 0) return;
  const int nValid = (int)nv;
  if (cta * perPass >= nValid) return;
  const bool relay = cta > 0;
  const int *progPrev = prog + 1 + (size_t)s * ctas + (cta > 0 ? cta - 1 : 0);
  int *progMine = prog + 1 + (size_t)s * ctas + cta;
  cf *stOut = st.outSpec + (size_t)s * C * B;
  const size_t CB = (size_t)C * B;
  const cf *specRot = T.specRot;
  const int handSrc = warp > 0 ? warp - 1 : 0;
  // SPLIT: the S5 lane of (block, channel) reads the previous block's output of that channel out of ITS ring: CPL lanes below, or
  // the last block of the warp below
  const cf *ringOPrev = bl > 0 ? ringO - CPL : (warp > 0 ? ringO - (size_t)(RN + RO) * 32 + (32 - CPL) : ringO);
  for (int i = tid; i < nW * (RN + RO) * 32; i += nThreads) { cf z; z.re = z.im = 0.f; rings[i] = z; }
  for (int i = tid; i < 2 * C * TL + 2 * nW * C; i += nThreads) { cf z; z.re = z.im = 0.f; tile[i] = z; }
  if (lane == 0) { for (int i = 0; i < NSTG; ++i) mbar_init(bars + i, 1); mbar_fence_init(); }
  __syncthreads();
  unsigned nIssued = 0, nWaited = 0;

  for (int p0 = cta * perPass; p0 < nValid; p0 += perPass * ctas) {
    const int slot = p0 + jb;
    const bool active = slot < nValid;
    const int lastJ = min(perPass - 1, nValid - 1 - p0);
    const bool isNew = active && (blocks[sd.blockBase + slot0 + slot].flags & kNew);
    const bool isLast = (jb == lastJ) && (ctas == 1 || p0 + lastJ == nValid - 1);   // writes the carried state
    const bool publishes = (jb == lastJ) && ctas > 1;
    const size_t blk = (size_t)s * nSlots + (active ? slot : p0);
    cf *so = specOut + blk * CB + (size_t)c * B;             // this lane's channel of its block's output spectrum
    const int tEnd = (B - 1 + ls) + lastJ * D;
    // record rows of this warp: its BPW blocks are lanes l0 .. l0+BPW-1 of record group gIdx; diagonal u of that group holds row
    // u - l*D of lane l, and the warp's piece of it is contiguous
    const int w0 = p0 + warp * BPW, gIdx = w0 >> 5, l0 = w0 & 31, uOff = (32 * gIdx - p0) * D;
    const bool warpLive = w0 < nValid;
    const float *grpBase = st.rec + ((size_t)s * ((nSlots + 31) / 32) + gIdx) * rec_group_floats(B, ls, C) + (size_t)l0 * NRP;
    float *stage = stageAll + (size_t)warp * NSTG * BPW * NRP;
    const int nDiag = rows + 31 * D;
    auto has_rows = [&](int t) { const int u = t + OA - uOff; return warpLive && u >= 0 && u < nDiag; };
    auto fetch = [&](int t) {
      if (has_rows(t) && t <= tEnd) {   // (nothing is fetched that no step will wait for)
        if (lane == 0 && (SPLIT ? roleS5 : true)) {   // (the S5 twin has the slack: issuing a copy costs some 400 cycles)
          const int u = t + OA - uOff;
          unsigned long long *bar = bars + (nIssued % NSTG);
          mbar_expect_tx(bar, BPW * NRP * 4);
          bulk_g2s(stage + (size_t)(nIssued % NSTG) * BPW * NRP, grpBase + (size_t)u * (32 * NRP), BPW * NRP * 4, bar);
        }
        ++nIssued;
      }
    };

    auto request_tile = [&](int ti) -> bool {   // as in chain_kernel: carried state, or the relayed predecessor's output
      const int b0 = ti * TL;
      if (b0 >= B) return true;
      cf *dst = tile + (size_t)(ti & 1) * C * TL;
      const cf *src = stOut;
      if (relay) {
        src = specOut + ((size_t)s * nSlots + p0 - 1) * CB;
        const bool ok = tid != 0 || relay_wait(progPrev, min(B, b0 + TL));
        if (__syncthreads_or(!ok)) {
          if (tid == 0) { atomicExch(err, 1); __threadfence(); st_release_gpu(progMine, kRelayPoison); }
          return false;
        }
      }
      for (int i = tid; i < C * (TL / 2); i += nThreads) {
        const int cc = i / (TL / 2), jj = (i - cc * (TL / 2)) * 2;
        if (b0 + jj < B) cp_async16(dst + (size_t)cc * TL + jj, src + (size_t)cc * B + b0 + jj);
      }
      return true;
    };
    // a CTA may not leave while bulk copies into its shared memory are in flight (the relay's time-out path)
    auto drain = [&]() { while (nWaited < nIssued) { mbar_wait(bars + (nWaited % NSTG), (nWaited / NSTG) & 1); ++nWaited; } };
    if (!request_tile(0)) { cp_async_commit(); cp_async_wait<0>(); return; }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    cf last; last.re = last.im = 0.f;
    for (int t = 0; t < NSTG - 1; ++t) fetch(t);             // the first stages
    cf rot; rot.re = rot.im = 0.f;
    { const int r = OA - jb * D; if (r >= 1 && r < B) rot = specRot[r]; }

    for (int t = 0; t <= tEnd; ++t) {
      if (!SPLIT) fetch(t + NSTG - 1);                       // its stage was last read in step t-1 (the __syncwarp below)
      cf rotNxt; rotNxt.re = rotNxt.im = 0.f;
      if (!SPLIT || roleS5) { const int r = t + 1 + OA - jb * D; if (r >= 1 && r < B) rotNxt = specRot[r]; }
      const int q0 = t + OA;
      if ((q0 % TL) == TL - kWideAhead) { const bool ok = request_tile(q0 / TL + 1); cp_async_commit(); if (!ok) { drain(); return; } }
      if ((q0 % TL) == TL - 1) cp_async_wait<0>();
      __syncthreads();
      if (SPLIT) fetch(t + NSTG - 1);                        // (both twins have read the stage of step t-1: the barrier)
      const int tau = t - jb * D, q = tau + OA, k = tau - ls;
      const bool validQ = active && q >= 1 && q < B, validK = active && k >= 0 && k < B;
      const float *row = stage + bl * NRP;
      if (has_rows(t)) {
        mbar_wait(bars + (nWaited % NSTG), (nWaited / NSTG) & 1);
        row += (size_t)(nWaited % NSTG) * BPW * NRP;
        ++nWaited;
      }
      if (SPLIT && roleS5) {
        // ---- S1 + S5 of this lane's channel at bin q, on its own warp: the previous block's output is in that block's ring (or the tile)
        if (__any_sync(full, validQ)) {
          const int qc = q & (2 * TL - 1);
          const cf oT = tile[((size_t)(qc / TL) * C + c) * TL + (qc % TL)];
          cf o = ringOPrev[(size_t)(q & RMO) * 32];
          if (bl == 0 && warp == 0) o = oT;
          const float tRe = row[SO + 3 * c], tIm = row[SO + 3 * c + 1], dv = row[SO + 3 * c + 2];
          bool slowQ = false;
          cf n5 = s5_fast(o, isNew, rot, tRe, tIm, dv, slowQ);
          if (validQ && slowQ) n5 = s5_bin(o, isNew, rot, tRe, tIm, dv);
          if (validQ) ringN[(size_t)(q & RMN) * 32] = n5;
        }
      } else if (__any_sync(full, SPLIT ? validK : (validQ || validK))) {
        if (!SPLIT) {
          // ---- S1 + S5 of this lane's channel at bin q: the previous block's output comes from CPL lanes below, the warp below, or the tile
          const int qc = q & (2 * TL - 1);
          cf o;
          o.re = __shfl_up_sync(full, last.re, CPL);
          o.im = __shfl_up_sync(full, last.im, CPL);
          const cf oT = tile[((size_t)(qc / TL) * C + c) * TL + (qc % TL)];
          const cf oH = hand[((size_t)((t & 1) ^ 1) * nW + handSrc) * C + c];
          if (bl == 0) o = (warp == 0) ? oT : oH;
          const float tRe = row[SO + 3 * c], tIm = row[SO + 3 * c + 1], dv = row[SO + 3 * c + 2];
          bool slowQ = false;
          cf n5 = s5_fast(o, isNew, rot, tRe, tIm, dv, slowQ);
          if (validQ && slowQ) n5 = s5_bin(o, isNew, rot, tRe, tIm, dv);
          if (validQ) ringN[(size_t)(q & RMN) * 32] = n5;    // (slot q = k + ls + 1: not one of the two read below)
        }
        // ---- S6 at bin k: the four neighbour values of the maximum channel, from its lane
        const int mc = validK ? __float_as_int(row[8]) : 0;
        const int src = bl * CPL + mc;
        const cf myLong = ringO[(size_t)((k - ls) & RMO) * 32], myN1 = ringN[(size_t)((k + 1) & RMN) * 32], myNL = ringN[(size_t)((k + ls) & RMN) * 32];
        cf oPrev, oLong, n1, nL;
        oPrev.re = __shfl_sync(full, last.re, src); oPrev.im = __shfl_sync(full, last.im, src);
        oLong.re = __shfl_sync(full, myLong.re, src); oLong.im = __shfl_sync(full, myLong.im, src);
        n1.re = __shfl_sync(full, myN1.re, src); n1.im = __shfl_sync(full, myN1.im, src);
        nL.re = __shfl_sync(full, myNL.re, src); nL.im = __shfl_sync(full, myNL.im, src);
        float ra[8];
        { const float4 a = *(const float4 *)row, b = *(const float4 *)(row + 4); ra[0] = a.x; ra[1] = a.y; ra[2] = a.z; ra[3] = a.w; ra[4] = b.x; ra[5] = b.y; ra[6] = b.z; ra[7] = b.w; }
        float phRe, phIm;
        chain_phase(ra, k, B, ls, oPrev, oLong, n1, nL, phRe, phIm);
        const float eMc = row[9 + 5 * mc]; cf fbMc; fbMc.re = row[9 + 5 * mc + 1]; fbMc.im = row[9 + 5 * mc + 2];
        const float eC = row[9 + 5 * c], wRe = row[9 + 5 * c + 3], wIm = row[9 + 5 * c + 4]; cf fbC; fbC.re = row[9 + 5 * c + 1]; fbC.im = row[9 + 5 * c + 2];
        bool slowK = false;
        cf om, out;
        make_output_fast(eMc, fbMc, phRe, phIm, om, slowK);
        {
          const float qIm = (wIm * om.re) + (wRe * om.im), qRe = (wRe * om.re) - (wIm * om.im);
          bool slowF = false;
          make_output_fast(eC, fbC, qRe, qIm, out, slowF);
          if (c == mc) out = om; else slowK |= slowF;
        }
        if (validK && slowK) {   // rare: the same with the plain IEEE operators
          make_output(eMc, fbMc, phRe, phIm, om.re, om.im);
          const float qIm = (wIm * om.re) + (wRe * om.im), qRe = (wRe * om.re) - (wIm * om.im);
          make_output(eC, fbC, qRe, qIm, out.re, out.im);
          if (c == mc) out = om;
        }
        if (g.incremental && active) {   // compat shim only: see compat_flush / BlockRec2::zeroBelow
          if (k < (int)blocks2[sd.blockBase + slot0 + slot].zeroBelow) out.re = out.im = 0.f;
        }
        if (validK) {
          ringO[(size_t)(k & RMO) * 32] = out;
          if (chOK) {
            if (k & 1) __stcs(reinterpret_cast<float4 *>(so + k - 1), make_float4(last.re, last.im, out.re, out.im));
            if (isLast) stOut[(size_t)c * B + k] = out;
            if (!SPLIT && bl == BPW - 1) hand[((size_t)(t & 1) * nW + warp) * C + c] = out;
          }
          last = out;
        }
        const bool pub = publishes && validK && ((k & (TL - 1)) == TL - 1 || k == B - 1);
        if (__any_sync(full, pub)) {    // relay: the last block's output up to bin k is in specOut, every channel of it: the warp barrier
          __syncwarp();                 // orders the channel lanes' stores before lane 0's release (which is cumulative)
          if (pub && cRaw == 0) st_release_gpu(progMine, k + 1);
        }
      }
      __syncwarp();                 // every lane has read its row of step t
      rot = rotNxt;
    }
    cp_async_wait<0>();
    __syncthreads();
  }
}

}  // namespace bs
