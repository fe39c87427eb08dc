// K8: frame-to-frame projection matching (sm_100a) and its C-ABI.
//
// Reference semantics (thirdparty/orb-slam2/src): ORBmatcher::SearchByProjection(Frame&,const Frame&,th,bMono=true)
// ORBmatcher.cc:1332-1474; Frame::GetFeaturesInArea / PosInGrid Frame.cc:331-396 (64x48 grid, candidate order =
// cell x ascending, cell y ascending, keypoint index ascending); DescriptorDistance ORBmatcher.cc:1651-1667;
// ComputeThreeMaxima :1605-1646; TH_HIGH=100, HISTO_LENGTH=30 (:38-40); retry at 2*th below 20 matches
// (Tracking.cc:876-883).
//
// One CTA per frame pair; the phases are described at k_match.  The reference's greedy loop is sequential in the
// query index -- a target taken by an earlier query is skipped -- which is why phase B replays the queries in order.
#include <cuda_runtime.h>

#include <vector>

#include "common.cuh"

namespace pgb {

constexpr int kMtThreads = 256;

constexpr int kMtK = 32;        // sorted candidates kept per query
constexpr int kThHigh = 100;
constexpr int kHisto = 30;
constexpr int kGridCols = 64, kGridRows = 48;

struct MatchArgs {
  int cap, nlevels, checkOri, consecutive, onlyIfBelow20;
  float minX, maxX, minY, maxY, th;
  float scale[16];
  // generic mode (consecutive == 0): arrays indexed [pair][cap]
  const pgb_keypoint* curK;
  const uint8_t* curD;
  const int* curN;
  const float* qUV;
  const int* qOct;
  const float* qAng;
  const uint8_t* qD;
  const uint8_t* qValid;
  const int* qN;
  // consecutive mode: frame arrays [frame][cap]; pair p = (frame p, frame p+1); flow[p][2]
  const float* flow;
  int* matchOfCur;   // [pair][cap]
  int* nMatches;     // [pair]
  uint32_t* qList32;          // scratch [pair][cap][kMtK]: target | distance << 16, in arrival order
  int* qCnt;                  // scratch [pair][cap] (total candidates found, may exceed kMtK)
  // work split: mode 1 = candidate search only (phases 0 + A) for the queries of split (blockIdx.x % nSplit) of pair
  // (blockIdx.x / nSplit), followed by k_match_resolve; mode 0 = the whole sequential kernel, run only for pairs
  // whose overflow flag is set (a query with more than kMtK candidates)
  int mode, nSplit;
  int* overflow;              // [pair]
};

__device__ __forceinline__ int hamming256(const uint32_t* a, const uint32_t* b) {
  int d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d += __popc(a[i] ^ b[i]);
  return d;
}

struct QueryWin {
  float u, v, r;
  int cx0, cx1, cy0, cy1, o0, o1;
  bool ok;
};

__device__ __forceinline__ QueryWin make_window(const MatchArgs& A, float u, float v, int octave, bool valid,
                                                float invW, float invH) {
  QueryWin q;
  q.u = u; q.v = v; q.ok = false; q.r = 0; q.cx0 = q.cx1 = q.cy0 = q.cy1 = 0;
  q.o0 = octave - 1; q.o1 = octave + 1;
  if (!valid) return q;
  if (u < A.minX || u > A.maxX || v < A.minY || v > A.maxY) return q;
  if (octave < 0 || octave >= A.nlevels) return q;
  q.r = A.th * A.scale[octave];
  const int nMinCellX = max(0, (int)floorf((u - A.minX - q.r) * invW));
  if (nMinCellX >= kGridCols) return q;
  const int nMaxCellX = min(kGridCols - 1, (int)ceilf((u - A.minX + q.r) * invW));
  if (nMaxCellX < 0) return q;
  const int nMinCellY = max(0, (int)floorf((v - A.minY - q.r) * invH));
  if (nMinCellY >= kGridRows) return q;
  const int nMaxCellY = min(kGridRows - 1, (int)ceilf((v - A.minY + q.r) * invH));
  if (nMaxCellY < 0) return q;
  q.cx0 = nMinCellX; q.cx1 = nMaxCellX; q.cy0 = nMinCellY; q.cy1 = nMaxCellY;
  q.ok = true;
  return q;
}

// shared-memory layout per CTA (dynamic): for each current keypoint 8 x u32 descriptor, x, y, angle (float),
// meta = posX | posY<<8 | octave<<16 | ingrid<<24, bin/taken word, the grid-cell-sorted index list; then the
// 64x48 grid's cell start offsets (u16).
//
// Phase 0: current keypoints are binned into the Frame grid (counting sort that keeps index order inside a cell,
//          Frame.cc:234-249).
// Phase A: one THREAD per query walks its window's cells in the reference's order (ix, iy, index ascending --
//          Frame.cc:352-381), applies the octave and |dx|,|dy| < r tests, takes the Hamming distance with POPC and
//          appends (distance, target) to the query's candidate row in arrival order.
// Phase B: one warp replays the reference's greedy loop in query order: lanes hold the query's candidates, taken
//          targets are masked out and a single warp-wide min over (distance << 8 | arrival) picks the winner --
//          strict '<' with first-wins ties, ORBmatcher.cc:1420-1433.  The next query's row is prefetched.
constexpr int kCells = kGridCols * kGridRows;  // 3072

__device__ __forceinline__ bool query_window(const MatchArgs& A, int p, int i, const pgb_keypoint* prevK, float fx,
                                             float fy, float invW, float invH, QueryWin* q) {
  float u, v;
  int oct;
  bool valid;
  if (A.consecutive) {
    const pgb_keypoint k = prevK[i];
    u = k.x + fx; v = k.y + fy; oct = k.octave; valid = true;
  } else {
    const size_t o = (size_t)p * A.cap + i;
    u = A.qUV[o * 2]; v = A.qUV[o * 2 + 1]; oct = A.qOct[o]; valid = A.qValid[o] != 0;
  }
  *q = make_window(A, u, v, oct, valid, invW, invH);
  return q->ok;
}

__global__ void __launch_bounds__(kMtThreads) k_match(MatchArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int p = A.mode == 1 ? blockIdx.x / A.nSplit : blockIdx.x, split = A.mode == 1 ? blockIdx.x % A.nSplit : 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cap = A.cap;
  if (A.onlyIfBelow20 && A.nMatches[p] >= 20) return;
  if (A.mode == 0 && A.overflow && !A.overflow[p]) return;

  const pgb_keypoint* curK;
  const uint8_t* curD;
  const pgb_keypoint* prevK = nullptr;
  const uint8_t* qD;
  int nCur, nQ;
  float fx = 0.f, fy = 0.f;
  if (A.consecutive) {
    curK = A.curK + (size_t)(p + 1) * cap;
    curD = A.curD + (size_t)(p + 1) * cap * 32;
    prevK = A.curK + (size_t)p * cap;
    qD = A.curD + (size_t)p * cap * 32;
    nCur = A.curN[p + 1];
    nQ = A.curN[p];
    fx = A.flow[2 * p]; fy = A.flow[2 * p + 1];
  } else {
    curK = A.curK + (size_t)p * cap;
    curD = A.curD + (size_t)p * cap * 32;
    qD = A.qD + (size_t)p * cap * 32;
    nCur = A.curN[p];
    nQ = A.qN[p];
  }
  nCur = min(nCur, cap);
  nQ = min(nQ, cap);

  uint32_t* sDesc = reinterpret_cast<uint32_t*>(smem);            // [cap][8]
  float* sX = reinterpret_cast<float*>(sDesc + (size_t)cap * 8);  // [cap]
  float* sY = sX + cap;
  uint32_t* sMeta = reinterpret_cast<uint32_t*>(sY + cap);        // [cap]
  float* sAng = reinterpret_cast<float*>(sMeta + cap);            // [cap]
  int* sBin = reinterpret_cast<int*>(sAng + cap);                 // [cap] -1 free, else histogram bin
  uint16_t* sOrder = reinterpret_cast<uint16_t*>(sBin + cap);     // [cap] target indices sorted by cell
  uint16_t* sCell = sOrder + cap + (cap & 1);                     // [kCells + 1] start offsets, then cursors
  uint16_t* sCur = sCell + kCells + 2;                            // [kCells] running cursors for the scatter
  float* sQAng = reinterpret_cast<float*>(sCur + kCells);         // [cap] query angles (phase B must not touch HBM)
  int* sMatch = reinterpret_cast<int*>(sQAng + cap);              // [cap] match_of_cur, written out at the end
  __shared__ int sHist[kHisto];
  __shared__ int sKeep[3];
  __shared__ int sNm;

  const float invW = (float)kGridCols / (A.maxX - A.minX);
  const float invH = (float)kGridRows / (A.maxY - A.minY);

  int* matchOfCur = A.matchOfCur + (size_t)p * cap;
  for (int i = tid; i < kCells + 1; i += kMtThreads) sCell[i] = 0;
  __syncthreads();
  for (int i = tid; i < nCur; i += kMtThreads) {
    const pgb_keypoint k = curK[i];
    sX[i] = k.x; sY[i] = k.y; sAng[i] = k.angle;
    const int posX = (int)roundf((k.x - A.minX) * invW), posY = (int)roundf((k.y - A.minY) * invH);
    const bool in = !(posX < 0 || posX >= kGridCols || posY < 0 || posY >= kGridRows);
    sMeta[i] = in ? ((uint32_t)posX | ((uint32_t)posY << 8) | ((uint32_t)(k.octave & 0xff) << 16) | (1u << 24)) : 0u;
    sBin[i] = -1;
    if (in) atomicAdd(reinterpret_cast<unsigned int*>(sCell) + ((posX * kGridRows + posY + 1) >> 1),
                      ((posX * kGridRows + posY + 1) & 1) ? 0x10000u : 1u);  // u16 histogram, two bins per word
  }
  for (int i = tid; i < cap; i += kMtThreads) sMatch[i] = -1;
  for (int i = tid; i < nCur * 8; i += kMtThreads) sDesc[i] = reinterpret_cast<const uint32_t*>(curD)[i];
  if (tid < kHisto) sHist[tid] = 0;
  if (tid == 0) sNm = 0;
  __syncthreads();
  // exclusive scan of the histogram (sCell[c + 1] holds the count of cell c): warp 0, 96 bins per lane
  if (warp == 0) {
    constexpr int per = kCells / 32;  // 96
    int sum = 0;
    for (int k = 0; k < per; k++) sum += sCell[1 + lane * per + k];
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    int run = incl - sum;
    for (int k = 0; k < per; k++) {
      const int c = sCell[1 + lane * per + k];
      sCur[lane * per + k] = (uint16_t)run;
      run += c;
      sCell[1 + lane * per + k] = (uint16_t)run;  // becomes the END offset of cell (lane*per + k) = start of the next
    }
  }
  __syncthreads();
  // scatter in index order (one warp, chunks of 32 in order; lanes of a chunk that share a cell are ranked by lane)
  if (warp == 0) {
    for (int t0 = 0; t0 < nCur; t0 += 32) {
      const int t = t0 + lane;
      const uint32_t m = t < nCur ? sMeta[t] : 0u;
      const bool in = (m >> 24) != 0;
      const int cell = in ? (int)(m & 0xff) * kGridRows + (int)((m >> 8) & 0xff) : -1 - lane;
      const uint32_t same = __match_any_sync(0xffffffffu, cell);
      if (in) {
        const int rank = __popc(same & ((1u << lane) - 1u));
        sOrder[sCur[cell] + rank] = (uint16_t)t;
      }
      __syncwarp();
      if (in && (same & ((1u << lane) - 1u)) == 0) sCur[cell] = (uint16_t)(sCur[cell] + __popc(same));
      __syncwarp();
    }
  }
  __syncthreads();

  uint32_t* qRow = A.qList32 + (size_t)p * cap * kMtK;  // [query][kMtK]: target | dist << 16, arrival order
  int* qCnt = A.qCnt + (size_t)p * cap;

  // ---------------- phase A: thread per query
  for (int i = tid + split * kMtThreads; i < nQ; i += kMtThreads * A.nSplit) {
    QueryWin q;
    int cnt = 0;
    sQAng[i] = A.consecutive ? prevK[i].angle : A.qAng[(size_t)p * cap + i];
    if (query_window(A, p, i, prevK, fx, fy, invW, invH, &q)) {
      uint32_t qd[8];
#pragma unroll
      for (int w = 0; w < 8; w++) qd[w] = reinterpret_cast<const uint32_t*>(qD)[(size_t)i * 8 + w];
      for (int ix = q.cx0; ix <= q.cx1; ix++)
        for (int iy = q.cy0; iy <= q.cy1; iy++) {
          const int cell = ix * kGridRows + iy;
          const int e1 = sCell[cell + 1];
          for (int e = cell ? sCell[cell] : 0; e < e1; e++) {
            const int t = sOrder[e];
            const int o = (sMeta[t] >> 16) & 0xff;
            if (o < q.o0 || o > q.o1) continue;
            if (!(fabsf(sX[t] - q.u) < q.r && fabsf(sY[t] - q.v) < q.r)) continue;
            const int d = hamming256(qd, sDesc + (size_t)t * 8);
            if (cnt < kMtK) qRow[(size_t)i * kMtK + cnt] = (uint32_t)t | ((uint32_t)d << 16);
            cnt++;
          }
        }
    }
    qCnt[i] = cnt;
  }
  if (A.mode == 1) return;  // the greedy assignment is resolved by k_match_resolve
  __syncthreads();

  // ---------------- phase B: greedy replay in query order
  if (warp == 0) {
    const float factor = 1.0f / kHisto;
    int nm = 0;
    // candidate rows are streamed through registers in blocks of kPB queries, one block ahead of the replay, so
    // the global-memory latency of a row is hidden behind the resolution of the previous block
    constexpr int kPB = 8;
    uint32_t rowN[kPB], row[kPB];
    int cN, cC = 0;
    auto load_block = [&](int i0, uint32_t (&r)[kPB], int& c) {
      c = (lane < kPB && i0 + lane < nQ) ? qCnt[i0 + lane] : 0;
#pragma unroll
      for (int k = 0; k < kPB; k++) r[k] = (i0 + k < nQ) ? qRow[(size_t)(i0 + k) * kMtK + lane] : 0u;
    };
    load_block(0, rowN, cN);
    for (int i = 0; i < nQ; i++) {
      const int k = i & (kPB - 1);
      if (k == 0) {
#pragma unroll
        for (int q = 0; q < kPB; q++) row[q] = rowN[q];
        cC = cN;
        load_block(i + kPB, rowN, cN);
      }
      const int cnt = __shfl_sync(0xffffffffu, cC, k);
      uint32_t ent = row[0];
#pragma unroll
      for (int q = 1; q < kPB; q++)
        if (k == q) ent = row[q];
      if (cnt == 0) continue;
      int winT = -1, winD = 256;
      if (cnt <= kMtK) {
        uint32_t key = 0xffffffffu;
        if (lane < cnt && sBin[ent & 0xffff] == -1) key = ((ent >> 16) << 8) | (uint32_t)lane;
        const uint32_t best = __reduce_min_sync(0xffffffffu, key);
        if (best != 0xffffffffu) {
          winD = (int)(best >> 8);
          winT = (int)(__shfl_sync(0xffffffffu, ent, best & 31) & 0xffff);
        }
      } else {
        // more candidates than a row holds: redo the window sweep for this query, skipping taken targets
        QueryWin q;
        query_window(A, p, i, prevK, fx, fy, invW, invH, &q);
        uint32_t qd[8];
#pragma unroll
        for (int w = 0; w < 8; w++) qd[w] = reinterpret_cast<const uint32_t*>(qD)[(size_t)i * 8 + w];
        unsigned long long best = ~0ull;
        for (int t = lane; t < nCur; t += 32) {
          const uint32_t m = sMeta[t];
          const int posX = m & 0xff, posY = (m >> 8) & 0xff, o = (m >> 16) & 0xff;
          if ((m >> 24) && sBin[t] == -1 && posX >= q.cx0 && posX <= q.cx1 && posY >= q.cy0 && posY <= q.cy1 &&
              o >= q.o0 && o <= q.o1 && fabsf(sX[t] - q.u) < q.r && fabsf(sY[t] - q.v) < q.r) {
            const int d = hamming256(qd, sDesc + (size_t)t * 8);
            const unsigned long long k2 =
                ((unsigned long long)d << 40) | ((unsigned long long)(posX * kGridRows + posY) << 20) | (unsigned)t;
            best = min(best, k2);
          }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, s));
        if (best != ~0ull) { winD = (int)(best >> 40); winT = (int)(best & 0xfffff); }
      }
      if (winT >= 0 && winD <= kThHigh) {
        int bin = kHisto;  // "assigned, no histogram"
        if (A.checkOri) {
          float rot = sQAng[i] - sAng[winT];
          if (rot < 0.0f) rot += 360.0f;
          bin = (int)roundf(rot * factor);
          if (bin == kHisto) bin = 0;
        }
        if (lane == 0) {
          sBin[winT] = bin;
          sMatch[winT] = i;
          if (A.checkOri) sHist[bin]++;
        }
        nm++;
      }
      __syncwarp();
    }
    if (lane == 0) {
      sNm = nm;
      int ind1 = -1, ind2 = -1, ind3 = -1;
      if (A.checkOri) {
        int max1 = 0, max2 = 0, max3 = 0;
        for (int b = 0; b < kHisto; b++) {
          const int s = sHist[b];
          if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = b; }
          else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = b; }
          else if (s > max3) { max3 = s; ind3 = b; }
        }
        if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
      }
      sKeep[0] = ind1; sKeep[1] = ind2; sKeep[2] = ind3;
    }
  }
  __syncthreads();
  if (A.checkOri) {
    int removed = 0;
    for (int t = tid; t < nCur; t += kMtThreads) {
      const int b = sBin[t];
      if (b >= 0 && b < kHisto && b != sKeep[0] && b != sKeep[1] && b != sKeep[2]) {
        sMatch[t] = -1;
        removed++;
      }
    }
    if (removed) atomicSub(&sNm, removed);
    __syncthreads();
  }
  for (int i = tid; i < cap; i += kMtThreads) matchOfCur[i] = sMatch[i];
  if (tid == 0) A.nMatches[p] = sNm;
}

// Greedy assignment of ORBmatcher.cc:1380-1447 without the sequential replay.  The reference visits the queries in
// index order and gives each its best still-free target.  Query i's decision is FINAL as soon as no unresolved
// query with a lower index has i's chosen target among its candidates (then nothing processed before i in the
// reference's order can still take it); targets taken by finalised queries of higher index can never be candidates of
// an unresolved lower one by the same rule.  Rounds: (1) every unresolved query atomicMin's its index into each of
// its free candidate targets, (2) every unresolved query picks its best free candidate (distance, then arrival
// order: strict '<', first wins) and finalises it if it holds the target's minimum.  The lowest unresolved query
// always finalises, chains of queries competing for the same targets resolve one link per round (a handful of
// rounds on real frames instead of ~1000 dependent iterations).  Identical results to the sequential replay.
__global__ void __launch_bounds__(kMtThreads) k_match_resolve(MatchArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int p = blockIdx.x, tid = threadIdx.x;
  const int cap = A.cap;
  if (A.onlyIfBelow20 && A.nMatches[p] >= 20) return;
  const pgb_keypoint* curK;
  const pgb_keypoint* prevK = nullptr;
  int nCur, nQ;
  if (A.consecutive) {
    curK = A.curK + (size_t)(p + 1) * cap;
    prevK = A.curK + (size_t)p * cap;
    nCur = A.curN[p + 1];
    nQ = A.curN[p];
  } else {
    curK = A.curK + (size_t)p * cap;
    nCur = A.curN[p];
    nQ = A.qN[p];
  }
  nCur = min(nCur, cap);
  nQ = min(nQ, cap);
  int* sBin = reinterpret_cast<int*>(smem);          // [cap] -1 free, else histogram bin (kHisto: taken, no histogram)
  int* sMinUn = sBin + cap;                          // [cap] lowest unresolved query listing the target
  int* sMatch = sMinUn + cap;                        // [cap]
  unsigned char* sRes = reinterpret_cast<unsigned char*>(sMatch + cap);  // [cap] query resolved
  __shared__ int sHist[kHisto];
  __shared__ int sKeep[3];
  __shared__ int sNm;
  const uint32_t* qRow = A.qList32 + (size_t)p * cap * kMtK;
  const int* qCnt = A.qCnt + (size_t)p * cap;

  int over = 0;
  for (int i = tid; i < cap; i += kMtThreads) {
    sBin[i] = -1; sMinUn[i] = 0x7fffffff; sMatch[i] = -1;
    sRes[i] = i < nQ ? 0 : 1;
    if (i < nQ && qCnt[i] > kMtK) over = 1;
  }
  if (tid < kHisto) sHist[tid] = 0;
  if (tid == 0) sNm = 0;
  over = __syncthreads_or(over);
  if (tid == 0) A.overflow[p] = over;
  if (over) return;  // a query's candidate row was truncated: the sequential kernel redoes this pair

  const float factor = 1.0f / kHisto;
  for (;;) {
    for (int i = tid; i < nQ; i += kMtThreads) {
      if (sRes[i]) continue;
      const int cnt = qCnt[i];
      for (int c = 0; c < cnt; c++) {
        const int t = (int)(qRow[(size_t)i * kMtK + c] & 0xffffu);
        if (sBin[t] == -1) atomicMin(&sMinUn[t], i);
      }
    }
    __syncthreads();
    int pending = 0;
    for (int i = tid; i < nQ; i += kMtThreads) {
      if (sRes[i]) continue;
      const int cnt = qCnt[i];
      uint32_t best = 0xffffffffu;
      int bestT = -1;
      for (int c = 0; c < cnt; c++) {
        const uint32_t e = qRow[(size_t)i * kMtK + c];
        const int t = (int)(e & 0xffffu);
        if (*(volatile int*)&sBin[t] != -1) continue;
        const uint32_t key = ((e >> 16) << 8) | (uint32_t)c;
        if (key < best) { best = key; bestT = t; }
      }
      if (bestT < 0 || (int)(best >> 8) > kThHigh) { sRes[i] = 1; continue; }  // no free candidate / bestDist > TH_HIGH
      if (sMinUn[bestT] != i) { pending = 1; continue; }
      int bin = kHisto;
      if (A.checkOri) {
        float rot = (A.consecutive ? prevK[i].angle : A.qAng[(size_t)p * cap + i]) - curK[bestT].angle;
        if (rot < 0.0f) rot += 360.0f;
        bin = (int)roundf(rot * factor);
        if (bin == kHisto) bin = 0;
        atomicAdd(&sHist[bin], 1);
      }
      *(volatile int*)&sBin[bestT] = bin;
      sMatch[bestT] = i;
      atomicAdd(&sNm, 1);
      sRes[i] = 1;
    }
    if (!__syncthreads_or(pending)) break;
    for (int t = tid; t < nCur; t += kMtThreads) sMinUn[t] = 0x7fffffff;
    __syncthreads();
  }
  if (tid == 0) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    if (A.checkOri) {  // ComputeThreeMaxima, ORBmatcher.cc:1605-1646
      int max1 = 0, max2 = 0, max3 = 0;
      for (int b = 0; b < kHisto; b++) {
        const int s = sHist[b];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = b; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = b; }
        else if (s > max3) { max3 = s; ind3 = b; }
      }
      if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
    }
    sKeep[0] = ind1; sKeep[1] = ind2; sKeep[2] = ind3;
  }
  __syncthreads();
  if (A.checkOri) {
    int removed = 0;
    for (int t = tid; t < nCur; t += kMtThreads) {
      const int b = sBin[t];
      if (b >= 0 && b < kHisto && b != sKeep[0] && b != sKeep[1] && b != sKeep[2]) {
        sMatch[t] = -1;
        removed++;
      }
    }
    if (removed) atomicSub(&sNm, removed);
    __syncthreads();
  }
  int* matchOfCur = A.matchOfCur + (size_t)p * cap;
  for (int i = tid; i < cap; i += kMtThreads) matchOfCur[i] = sMatch[i];
  if (tid == 0) A.nMatches[p] = sNm;
}

// ------------------------------------------------------------------------------------------------------------------
// The other windowed searches of ORBmatcher on the same primitive.  Both are sequential in
// the query index with state carried between queries, and neither is on the per-frame path of the benchmark, so one
// CTA per problem does: (A) thread per query: every target inside the query's GetFeaturesInArea window (Frame.cc:331-384:
// cell range, level range, |dx| < r and |dy| < r) with its Hamming distance, appended to the query's row; (B) thread 0
// replays the reference's loop over the rows.  A candidate's arrival order in the reference (cell x, cell y, index)
// only matters for ties of the strict '<' comparisons, so (B) compares (distance, cell, index) keys.
//   flavour 1: SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)      ORBmatcher.cc:407-522
//   flavour 2: SearchByProjection(Frame&, const vector<MapPoint*>&, th)                      ORBmatcher.cc:46-131
constexpr int kWinK = 96;       // candidates kept per query; more raises PGB_ERR_CAPACITY (never truncated silently)
constexpr int kThLow = 50;

struct WinArgs {
  int flavour, cap, nlevels, checkOri;
  float minX, maxX, minY, maxY, th, nnratio;
  float scale[16];
  const pgb_keypoint* curK;   // [prob][cap] targets (F2 / the frame)
  const uint8_t* curD;
  const int* curN;
  const uint8_t* curTaken;    // flavour 2: the feature already holds a map point with observations
  const float* qUV;           // [prob][cap][2]   flavour 1: vbPrevMatched, flavour 2: mTrackProjX/Y
  const int* qLevel;          // flavour 1: octave of the F1 keypoint, flavour 2: mnTrackScaleLevel
  const float* qAux;          // flavour 1: angle of the F1 keypoint, flavour 2: mTrackViewCos
  const uint8_t* qD;
  const uint8_t* qValid;      // flavour 2: mbTrackInView && !isBad()
  const uint8_t* qObs;        // flavour 2: Observations() > 0 of the map point
  const int* qN;
  int* matchOfQ;              // flavour 1: vnMatches12
  int* matchOfCur;            // flavour 2: index of the map point assigned to each feature by THIS call, else -1
  float* qUVOut;              // flavour 1: updated vbPrevMatched
  int* nMatches;
  uint32_t* rows;             // scratch [prob][cap][kWinK]: target | dist << 16
  int* rowCnt;                // scratch [prob][cap]
  int* err;
};

__global__ void __launch_bounds__(kMtThreads) k_match_windowed(WinArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int p = blockIdx.x, tid = threadIdx.x, cap = A.cap;
  const pgb_keypoint* curK = A.curK + (size_t)p * cap;
  const uint32_t* curD = reinterpret_cast<const uint32_t*>(A.curD + (size_t)p * cap * 32);
  const uint32_t* qD = reinterpret_cast<const uint32_t*>(A.qD + (size_t)p * cap * 32);
  const int nCur = min(A.curN[p], cap), nQ = min(A.qN[p], cap);
  uint16_t* sCellOf = reinterpret_cast<uint16_t*>(smem);       // [cap] cell (x*48+y) of each target, 0xffff = outside
  int* sState = reinterpret_cast<int*>(sCellOf + cap + (cap & 1));  // [cap] f1: vMatchedDistance; f2: taken flag
  int* sM21 = sState + cap;                                    // [cap] f1: vnMatches21; f2: match_of_cur
  int* sM12 = sM21 + cap;                                      // [cap] f1: vnMatches12
  signed char* sBin = reinterpret_cast<signed char*>(sM12 + cap);  // [cap] f1: histogram bin the query was pushed to
  const float invW = (float)kGridCols / (A.maxX - A.minX), invH = (float)kGridRows / (A.maxY - A.minY);
  uint32_t* rows = A.rows + (size_t)p * cap * kWinK;
  int* rowCnt = A.rowCnt + (size_t)p * cap;

  for (int t = tid; t < cap; t += kMtThreads) {
    uint16_t cell = 0xffff;
    if (t < nCur) {
      const pgb_keypoint k = curK[t];
      const int posX = (int)roundf((k.x - A.minX) * invW), posY = (int)roundf((k.y - A.minY) * invH);  // Frame::PosInGrid
      if (!(posX < 0 || posX >= kGridCols || posY < 0 || posY >= kGridRows)) cell = (uint16_t)(posX * kGridRows + posY);
    }
    sCellOf[t] = cell;
    sState[t] = A.flavour == 1 ? 0x7fffffff : (A.curTaken && t < nCur ? A.curTaken[(size_t)p * cap + t] : 0);
    sM21[t] = -1; sM12[t] = -1; sBin[t] = -1;
  }
  __syncthreads();

  // ---- (A) candidate rows
  for (int i = tid; i < nQ; i += kMtThreads) {
    const size_t o = (size_t)p * cap + i;
    int cnt = 0;
    const int level = A.qLevel[o];
    bool ok = A.flavour == 1 ? (level <= 0) : (A.qValid[o] != 0);          // "if(level1>0) continue" / mbTrackInView, isBad
    float r = A.th;                                                        // flavour 1: windowSize
    int o0 = 0, o1 = 0;
    if (A.flavour == 2) {
      if (level < 0 || level >= A.nlevels) ok = false;
      float rr = (double)A.qAux[o] > 0.998 ? 2.5f : 4.0f;                   // RadiusByViewingCos: float vs the DOUBLE literal (ORBmatcher.cc:133-139)
      if (A.th != 1.0f) rr *= A.th;
      r = ok ? rr * A.scale[level] : 0.f;
      o0 = level - 1; o1 = level;
    }
    if (ok) {
      const float u = A.qUV[2 * o], v = A.qUV[2 * o + 1];
      // GetFeaturesInArea's cell range with its four early exits (Frame.cc:336-350)
      const int cx0 = max(0, (int)floorf((u - A.minX - r) * invW)), cx1 = min(kGridCols - 1, (int)ceilf((u - A.minX + r) * invW));
      const int cy0 = max(0, (int)floorf((v - A.minY - r) * invH)), cy1 = min(kGridRows - 1, (int)ceilf((v - A.minY + r) * invH));
      if (!(cx0 >= kGridCols || cx1 < 0 || cy0 >= kGridRows || cy1 < 0)) {
        uint32_t qd[8];
#pragma unroll
        for (int w = 0; w < 8; w++) qd[w] = qD[(size_t)i * 8 + w];
        for (int t = 0; t < nCur; t++) {
          const int cell = sCellOf[t];
          if (cell == 0xffff) continue;
          const int px = cell / kGridRows, py = cell - px * kGridRows;
          if (px < cx0 || px > cx1 || py < cy0 || py > cy1) continue;
          const pgb_keypoint k = curK[t];
          if (k.octave < o0 || k.octave > o1) continue;
          if (!(fabsf(k.x - u) < r && fabsf(k.y - v) < r)) continue;
          const int d = hamming256(qd, curD + (size_t)t * 8);
          if (cnt < kWinK) rows[(size_t)i * kWinK + cnt] = (uint32_t)t | ((uint32_t)d << 16);
          cnt++;
        }
      }
    }
    rowCnt[i] = cnt;
    if (cnt > kWinK) atomicOr(A.err, 1);
  }
  __syncthreads();
  if (tid != 0) return;

  // ---- (B) sequential replay
  int nm = 0;
  auto key_of = [&](uint32_t e) -> unsigned long long {  // (distance, arrival order = cell then index)
    const int t = e & 0xffff;
    return ((unsigned long long)(e >> 16) << 40) | ((unsigned long long)sCellOf[t] << 20) | (unsigned)t;
  };
  if (A.flavour == 1) {
    int hist[kHisto];
    for (int b = 0; b < kHisto; b++) hist[b] = 0;
    const float factor = 1.0f / kHisto;
    for (int i = 0; i < nQ; i++) {
      const int cnt = min(rowCnt[i], kWinK);
      unsigned long long best = ~0ull, best2 = ~0ull;
      for (int c = 0; c < cnt; c++) {
        const uint32_t e = rows[(size_t)i * kWinK + c];
        if (sState[e & 0xffff] <= (int)(e >> 16)) continue;            // vMatchedDistance[i2] <= dist
        const unsigned long long k = key_of(e);
        if (k < best) { best2 = best; best = k; } else if (k < best2) best2 = k;
      }
      if (best == ~0ull) continue;
      const int bestDist = (int)(best >> 40), bestIdx2 = (int)(best & 0xfffff);
      const int bestDist2 = best2 == ~0ull ? 0x7fffffff : (int)(best2 >> 40);
      if (bestDist <= kThLow && (float)bestDist < (float)bestDist2 * A.nnratio) {
        if (sM21[bestIdx2] >= 0) { sM12[sM21[bestIdx2]] = -1; nm--; }
        sM12[i] = bestIdx2; sM21[bestIdx2] = i; sState[bestIdx2] = bestDist; nm++;
        if (A.checkOri) {
          float rot = A.qAux[(size_t)p * cap + i] - curK[bestIdx2].angle;
          if (rot < 0.0f) rot += 360.0f;
          int bin = (int)roundf(rot * factor);
          if (bin == kHisto) bin = 0;
          sBin[i] = (signed char)bin;   // a query is pushed to rotHist at most once
          hist[bin]++;
        }
      }
    }
    if (A.checkOri) {
      int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
      for (int b = 0; b < kHisto; b++) {
        const int s = hist[b];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = b; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = b; }
        else if (s > max3) { max3 = s; ind3 = b; }
      }
      if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
      for (int i = 0; i < nQ; i++) {
        const int b = sBin[i];
        if (b < 0 || b == ind1 || b == ind2 || b == ind3) continue;
        if (sM12[i] >= 0) { sM12[i] = -1; nm--; }
      }
    }
    for (int i = 0; i < cap; i++) {
      const size_t o = (size_t)p * cap + i;
      A.matchOfQ[o] = i < nQ ? sM12[i] : -1;
      if (i < nQ) {  // "Update prev matched"
        const bool m = sM12[i] >= 0;
        A.qUVOut[2 * o] = m ? curK[sM12[i]].x : A.qUV[2 * o];
        A.qUVOut[2 * o + 1] = m ? curK[sM12[i]].y : A.qUV[2 * o + 1];
      }
    }
  } else {
    for (int i = 0; i < nQ; i++) {
      const int cnt = min(rowCnt[i], kWinK);
      unsigned long long best = ~0ull, best2 = ~0ull;
      for (int c = 0; c < cnt; c++) {
        const uint32_t e = rows[(size_t)i * kWinK + c];
        if (sState[e & 0xffff]) continue;                               // F.mvpMapPoints[idx] with Observations() > 0
        const unsigned long long k = key_of(e);
        if (k < best) { best2 = best; best = k; } else if (k < best2) best2 = k;
      }
      if (best == ~0ull) continue;
      const int bestDist = (int)(best >> 40), bestIdx = (int)(best & 0xfffff);
      if (bestDist > kThHigh) continue;
      if (best2 != ~0ull) {
        const int bestDist2 = (int)(best2 >> 40), idx2 = (int)(best2 & 0xfffff);
        if (curK[bestIdx].octave == curK[idx2].octave && (float)bestDist > A.nnratio * (float)bestDist2) continue;
      }
      sM21[bestIdx] = i;
      sState[bestIdx] = A.qObs[(size_t)p * cap + i] ? 1 : 0;
      nm++;
    }
    for (int t = 0; t < cap; t++) A.matchOfCur[(size_t)p * cap + t] = t < nCur ? sM21[t] : -1;
  }
  A.nMatches[p] = nm;
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (ORBmatcher.cc:161-290).  The two DBoW2 feature vectors
// arrive as CSR arrays sorted by node id.  A frame feature belongs to exactly one vocabulary node, so the reference's
// only sequential dependence ("skip features of F that already got a map point") stays inside a node: one warp per
// common node replays that node's keyframe features in order, the lanes share the node's frame features, the best and
// second-best distances are two warp REDUX on (distance, position) keys.  One CTA per (keyframe, frame) problem; the
// rotation histogram lives in shared memory and is applied after a CTA barrier.
struct BowArgs {
  int cap, checkOri;
  float nnratio;
  const uint8_t* kfD;        // [prob][cap][32]
  const float* kfAng;        // [prob][cap]    mvKeysUn[i].angle
  const uint8_t* kfHas;      // [prob][cap]    vpMapPointsKF[i] && !isBad()
  const int* kfNodeOff;      // [prob + 1]     node range of each problem
  const uint32_t* kfNode;    // node ids, strictly increasing inside a problem
  const int* kfStart;        // [nodes + 1]    CSR over kfIdx
  const uint32_t* kfIdx;
  const uint8_t* fD;
  const float* fAng;
  const int* fN;             // [prob] F.N
  const int* fNodeOff;
  const uint32_t* fNode;
  const int* fStart;
  const uint32_t* fIdx;
  int* matchOfF;             // [prob][cap]
  int* nMatches;             // [prob]
  int* err;
};

__global__ void __launch_bounds__(kMtThreads) k_match_bow(BowArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cap = A.cap;
  int* sHist = reinterpret_cast<int*>(smem);                       // [32]: 30 bins, [30] = match count
  uint32_t* sSeen = reinterpret_cast<uint32_t*>(sHist + 32);       // [(cap + 31) / 32] frame feature listed in a node
  signed char* sBin = reinterpret_cast<signed char*>(sSeen + (cap + 31) / 32);  // [cap] bin the feature was pushed to
  int* mOf = A.matchOfF + (size_t)p * cap;
  const int nF = min(A.fN[p], cap);
  const int k0 = A.kfNodeOff[p], k1 = A.kfNodeOff[p + 1], f0 = A.fNodeOff[p], f1 = A.fNodeOff[p + 1];
  for (int t = tid; t < cap; t += kMtThreads) { mOf[t] = -1; sBin[t] = -1; }
  for (int t = tid; t < (cap + 31) / 32; t += kMtThreads) sSeen[t] = 0;
  if (tid < 32) sHist[tid] = 0;
  __syncthreads();

  // input contract: node ids strictly increasing, indices in range, a frame feature in at most one node
  bool bad = false;
  for (int k = k0 + tid; k + 1 < k1; k += kMtThreads) bad |= A.kfNode[k] >= A.kfNode[k + 1];
  for (int k = f0 + tid; k + 1 < f1; k += kMtThreads) bad |= A.fNode[k] >= A.fNode[k + 1];
  if (k1 > k0)
    for (int i = A.kfStart[k0] + tid; i < A.kfStart[k1]; i += kMtThreads) bad |= A.kfIdx[i] >= (uint32_t)cap;
  if (f1 > f0)
    for (int i = A.fStart[f0] + tid; i < A.fStart[f1]; i += kMtThreads) {
      const uint32_t x = A.fIdx[i];
      if (x >= (uint32_t)nF) { bad = true; continue; }
      if (atomicOr(&sSeen[x >> 5], 1u << (x & 31)) & (1u << (x & 31))) bad = true;
    }
  if (bad) atomicOr(A.err, 1);
  __syncthreads();

  const uint32_t* kfD = reinterpret_cast<const uint32_t*>(A.kfD + (size_t)p * cap * 32);
  const uint32_t* fD = reinterpret_cast<const uint32_t*>(A.fD + (size_t)p * cap * 32);
  const float factor = 1.0f / kHisto;
  for (int k = k0 + warp; k < k1; k += kMtThreads / 32) {
    const uint32_t id = A.kfNode[k];
    int lo = f0, hi = f1;                                   // lower_bound of id among F's nodes
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (A.fNode[mid] < id) lo = mid + 1; else hi = mid;
    }
    if (lo >= f1 || A.fNode[lo] != id) continue;
    const int fs = A.fStart[lo], fe = A.fStart[lo + 1];
    for (int iKF = A.kfStart[k]; iKF < A.kfStart[k + 1]; iKF++) {
      const uint32_t realIdxKF = A.kfIdx[iKF];
      if (realIdxKF >= (uint32_t)cap || !A.kfHas[(size_t)p * cap + realIdxKF]) continue;
      uint32_t q[8];
#pragma unroll
      for (int w = 0; w < 8; w++) q[w] = kfD[(size_t)realIdxKF * 8 + w];
      uint32_t b1 = 0xffffffffu, b2 = 0xffffffffu;         // (distance << 16 | position in the node), two smallest
      for (int iF = fs + lane; iF < fe; iF += 32) {
        const uint32_t realIdxF = A.fIdx[iF];
        if (realIdxF >= (uint32_t)nF || mOf[realIdxF] >= 0) continue;
        const uint32_t key = ((uint32_t)hamming256(q, fD + (size_t)realIdxF * 8) << 16) | (uint32_t)(iF - fs);
        if (key < b1) { b2 = b1; b1 = key; } else if (key < b2) b2 = key;
      }
      const uint32_t best = __reduce_min_sync(0xffffffffu, b1);
      const uint32_t second = __reduce_min_sync(0xffffffffu, b1 == best ? b2 : b1);
      if (lane == 0 && best != 0xffffffffu) {
        const int bestDist1 = (int)(best >> 16), bestDist2 = second == 0xffffffffu ? 256 : (int)(second >> 16);
        if (bestDist1 <= kThLow && (float)bestDist1 < A.nnratio * (float)bestDist2) {
          const uint32_t bestIdxF = A.fIdx[fs + (int)(best & 0xffffu)];
          mOf[bestIdxF] = (int)realIdxKF;
          if (A.checkOri) {
            float rot = A.kfAng[(size_t)p * cap + realIdxKF] - A.fAng[(size_t)p * cap + bestIdxF];
            if (rot < 0.0f) rot += 360.0f;
            int bin = (int)roundf(rot * factor);
            if (bin == kHisto) bin = 0;
            sBin[bestIdxF] = (signed char)bin;
            atomicAdd(&sHist[bin], 1);
          }
          atomicAdd(&sHist[30], 1);
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (A.checkOri) {
    int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;   // ComputeThreeMaxima, every thread the same
    for (int b = 0; b < kHisto; b++) {
      const int s = sHist[b];
      if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = b; }
      else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = b; }
      else if (s > max3) { max3 = s; ind3 = b; }
    }
    if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
    for (int t = tid; t < nF; t += kMtThreads) {
      const int b = sBin[t];
      if (b < 0 || b == ind1 || b == ind2 || b == ind3) continue;
      mOf[t] = -1;
      atomicSub(&sHist[30], 1);
    }
    __syncthreads();
  }
  if (tid == 0) A.nMatches[p] = sHist[30];
}

// MapPoint::ComputeDistinctiveDescriptors (thirdparty/orb-slam2/src/MapPoint.cc:259-324): among the N descriptors that
// observe a map point, the one with the least median Hamming distance to the rest (median = sorted row[(size_t)(0.5 *
// (N - 1))], the row includes the zero self-distance; first index wins ties).  One warp per map point; for each
// candidate i the lanes hold the distances to j = lane, lane + 32, ... and the k-th smallest is found by bisection on
// the value (distances are integers in [0, 256]) with ballot counts -- no sort.
constexpr int kDistinctMaxN = 256;

__global__ void __launch_bounds__(128) k_distinctive(const uint8_t* __restrict__ desc, const int* __restrict__ offsets,
                                                     int nPoints, int* __restrict__ bestIdx, int* __restrict__ err) {
  const int p = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (p >= nPoints) return;
  const int o = offsets[p], N = offsets[p + 1] - o;
  if (N <= 0) { if (lane == 0) bestIdx[p] = -1; return; }          // "if(vDescriptors.empty()) return;"
  if (N > kDistinctMaxN) { if (lane == 0) { bestIdx[p] = -1; atomicOr(err, 1); } return; }
  const uint32_t* D = reinterpret_cast<const uint32_t*>(desc) + (size_t)o * 8;
  const int k = (int)(0.5 * (N - 1));
  int best = 0x7fffffff, bestI = 0;
  for (int i = 0; i < N; i++) {
    uint32_t di[8];
#pragma unroll
    for (int w = 0; w < 8; w++) di[w] = D[(size_t)i * 8 + w];
    int d[kDistinctMaxN / 32];
#pragma unroll
    for (int c = 0; c < kDistinctMaxN / 32; c++) {
      const int j = lane + 32 * c;
      d[c] = j < N ? hamming256(di, D + (size_t)j * 8) : 0x7fffffff;
    }
    int lo = 0, hi = 256;  // smallest v with #(d <= v) >= k + 1
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      int cnt = 0;
#pragma unroll
      for (int c = 0; c < kDistinctMaxN / 32; c++) cnt += __popc(__ballot_sync(0xffffffffu, d[c] <= mid));
      if (cnt >= k + 1) hi = mid; else lo = mid + 1;
    }
    if (lo < best) { best = lo; bestI = i; }
  }
  if (lane == 0) bestIdx[p] = bestI;
}

size_t match_smem_bytes(int cap) {
  size_t b = (size_t)cap * 8 * 4 + (size_t)cap * 5 * 4;        // desc, x, y, meta, angle, bin
  b += (size_t)(cap + (cap & 1)) * 2;                           // cell-sorted order
  b += (size_t)(kCells + 2) * 2 + (size_t)kCells * 2;           // cell offsets + cursors
  b += (size_t)cap * 8;                                         // query angles, match_of_cur
  return b + 16;
}

__global__ void k_desc_distance(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int n,
                                int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = hamming256(reinterpret_cast<const uint32_t*>(a) + (size_t)i * 8, reinterpret_cast<const uint32_t*>(b) + (size_t)i * 8);
}

}  // namespace pgb

using namespace pgb;

struct pgb_matcher {
  int device = 0, checkOri = 1, maxFeats = 0, maxBatch = 0;
  float nnratio = 0.f;
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  DevBuf<uint32_t> qList;
  DevBuf<int> qCnt;
  // staging for host-buffer calls
  DevBuf<pgb_keypoint> dK;
  DevBuf<uint8_t> dD, dQD, dQV;
  DevBuf<float> dUV, dAng, dFlow;
  DevBuf<int> dN, dQN, dOct, dMatch, dNm;
  DevBuf<int> overflow;
};

namespace {

int launch_match(pgb_matcher* m, MatchArgs& A, int nPairs) {
  NvtxRange range(A.onlyIfBelow20 ? "pgb:match:retry_2th" : "pgb:match:search_by_projection");
  if (A.cap > m->maxFeats) return fail(PGB_ERR_CAPACITY, "cap %d exceeds the matcher's max_feats %d", A.cap, m->maxFeats);
  if (nPairs > m->maxBatch) return fail(PGB_ERR_CAPACITY, "n_pairs %d exceeds the matcher's max_batch %d", nPairs, m->maxBatch);
  const size_t smem = match_smem_bytes(A.cap);
  if (smem > 227 * 1024) return fail(PGB_ERR_CAPACITY, "cap %d needs %zu B of shared memory (max 227 KB)", A.cap, smem);
  static DynSmemLimit limMatch, limResolve;
  if (int rc = limMatch.ensure(k_match, smem)) return rc;
  A.qList32 = m->qList.p;
  A.qCnt = m->qCnt.p;
  A.overflow = m->overflow.p;
  // candidate search spread over kSplit CTAs per pair, then the round-based greedy resolution (one CTA per pair)
  constexpr int kSplit = 4;
  A.mode = 1; A.nSplit = kSplit;
  k_match<<<nPairs * kSplit, kMtThreads, smem, m->stream>>>(A);
  PGB_CHECK_LAUNCH();
  const size_t smem2 = (size_t)A.cap * 13 + 16;
  if (int rc = limResolve.ensure(k_match_resolve, smem2)) return rc;
  k_match_resolve<<<nPairs, kMtThreads, smem2, m->stream>>>(A);
  PGB_CHECK_LAUNCH();
  // pairs with a truncated candidate row (flagged by k_match_resolve) are redone by the sequential kernel
  A.mode = 0; A.nSplit = 1;
  k_match<<<nPairs, kMtThreads, smem, m->stream>>>(A);
  PGB_CHECK_LAUNCH();
  return PGB_OK;
}

void fill_common(MatchArgs& A, int cap, float minX, float maxX, float minY, float maxY, float th, const float* sf,
                 int nlevels, int checkOri) {
  memset(&A, 0, sizeof A);
  A.cap = cap; A.nlevels = nlevels; A.checkOri = checkOri;
  A.minX = minX; A.maxX = maxX; A.minY = minY; A.maxY = maxY; A.th = th;
  for (int i = 0; i < nlevels && i < 16; i++) A.scale[i] = sf[i];
}

}  // namespace

namespace pgb {
// Median displacement of the matched keypoints of a frame pair (prev = frame p, cur = frame p + 1 of the per-frame
// arrays): what the flow-tracking loop of optical_trajectories needs from a pair, computed where the matches are
// instead of after a device-to-host copy of every keypoint.  The value is the element at index n / 2 of the sorted
// displacements (std::nth_element in the host code this replaces), per axis: an element's rank is the number of elements
// that are smaller, or equal with a lower index -- one pass of n comparisons per element, no sort.  One CTA per pair.
constexpr int kFlowCap = 2048;
__global__ void __launch_bounds__(256) k_median_flow(int cap, const pgb_keypoint* __restrict__ kps, const int32_t* __restrict__ counts,
                                                     const int32_t* __restrict__ matchOfCur, const int32_t* __restrict__ nMatches,
                                                     float* __restrict__ flow, int32_t* __restrict__ tracked) {
  __shared__ float sx[kFlowCap], sy[kFlowCap];
  __shared__ int sn;
  const int p = blockIdx.x, tid = threadIdx.x;
  const pgb_keypoint* prevK = kps + (size_t)p * cap;
  const pgb_keypoint* curK = kps + (size_t)(p + 1) * cap;
  const int32_t* m = matchOfCur + (size_t)p * cap;
  const int nCur = min(counts[p + 1], cap);
  if (tid == 0) sn = 0;
  __syncthreads();
  // compaction in keypoint order is not needed: the median does not depend on the order of the elements
  for (int t = tid; t < nCur; t += blockDim.x) {
    const int q = m[t];
    if (q >= 0) {
      const int i = atomicAdd(&sn, 1);
      sx[i] = __fsub_rn(curK[t].x, prevK[q].x);
      sy[i] = __fsub_rn(curK[t].y, prevK[q].y);
    }
  }
  __syncthreads();
  const int n = sn, k = n / 2;
  const bool ok = nMatches[p] >= 20 && n > 0;
  if (tid == 0) tracked[p] = ok ? 1 : 0;
  if (!ok) {
    if (tid < 2) flow[2 * p + tid] = 0.f;
    return;
  }
  for (int i = tid; i < n; i += blockDim.x) {
    const float vx = sx[i], vy = sy[i];
    int rx = 0, ry = 0;
    for (int j = 0; j < n; j++) {
      const float ux = sx[j], uy = sy[j];
      rx += (ux < vx) || (ux == vx && j < i);
      ry += (uy < vy) || (uy == vy && j < i);
    }
    if (rx == k) flow[2 * p] = vx;      // ranks are a permutation: exactly one element per axis has rank k
    if (ry == k) flow[2 * p + 1] = vy;
  }
}
}  // namespace pgb

extern "C" {

pgb_matcher* pgb_matcher_create(int device, float nnratio, int check_orientation, int max_feats, int max_batch,
                                void* stream) {
  if (max_feats <= 0 || max_batch <= 0 || max_feats > (1 << 20) - 1) {
    fail(PGB_ERR_INVALID, "pgb_matcher_create: invalid argument");
    return nullptr;
  }
  if (use_device(device)) return nullptr;
  pgb_matcher* m = new pgb_matcher;
  m->device = device; m->nnratio = nnratio; m->checkOri = check_orientation ? 1 : 0;
  m->maxFeats = max_feats; m->maxBatch = max_batch;
  if (stream) m->stream = (cudaStream_t)stream;
  else {
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) {
      fail(PGB_ERR_CUDA, "cudaStreamCreate failed");
      delete m;
      return nullptr;
    }
    m->ownStream = true;
  }
  const size_t n = (size_t)max_feats * max_batch;
  if (m->qList.alloc(n * kMtK) || m->qCnt.alloc(n) || m->overflow.alloc(max_batch)) {
    pgb_matcher_destroy(m);
    return nullptr;
  }
  return m;
}

void pgb_matcher_destroy(pgb_matcher* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  if (m->ownStream && m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

int pgb_match_by_projection(pgb_matcher* m, int n_pairs, int cap, const pgb_keypoint* cur_kps, const uint8_t* cur_desc,
                            const int32_t* cur_counts, const float* q_uv, const int32_t* q_octave,
                            const float* q_angle, const uint8_t* q_desc, const uint8_t* q_valid,
                            const int32_t* q_counts, float min_x, float max_x, float min_y, float max_y, float th,
                            const float* scale_factors, int nlevels, int32_t* match_of_cur, int32_t* n_matches,
                            int is_device) {
  if (!m) return fail(PGB_ERR_INVALID, "null handle");
  if (n_pairs < 0 || cap <= 0 || nlevels <= 0 || nlevels > 16 || !scale_factors || !(max_x > min_x) || !(max_y > min_y))
    return fail(PGB_ERR_INVALID, "pgb_match_by_projection: invalid argument");
  if (n_pairs == 0) return PGB_OK;
  if (!cur_kps || !cur_desc || !cur_counts || !q_uv || !q_octave || !q_angle || !q_desc || !q_valid || !q_counts ||
      !match_of_cur || !n_matches)
    return fail(PGB_ERR_INVALID, "pgb_match_by_projection: null buffer");
  PGB_CUDA(cudaSetDevice(m->device));
  MatchArgs A;
  fill_common(A, cap, min_x, max_x, min_y, max_y, th, scale_factors, nlevels, m->checkOri);
  const size_t n = (size_t)n_pairs * cap;
  if (is_device) {
    A.curK = cur_kps; A.curD = cur_desc; A.curN = cur_counts; A.qUV = q_uv; A.qOct = q_octave; A.qAng = q_angle;
    A.qD = q_desc; A.qValid = q_valid; A.qN = q_counts; A.matchOfCur = match_of_cur; A.nMatches = n_matches;
    return launch_match(m, A, n_pairs);
  }
  if (m->dK.n < n) {
    if (m->dK.alloc(n) || m->dD.alloc(n * 32) || m->dQD.alloc(n * 32) || m->dQV.alloc(n) || m->dUV.alloc(n * 2) ||
        m->dAng.alloc(n) || m->dOct.alloc(n) || m->dMatch.alloc(n))
      return PGB_ERR_CUDA;
  }
  if (m->dN.n < (size_t)n_pairs) {
    if (m->dN.alloc(n_pairs) || m->dQN.alloc(n_pairs) || m->dNm.alloc(n_pairs)) return PGB_ERR_CUDA;
  }
  cudaStream_t s = m->stream;
  PGB_CUDA(cudaMemcpyAsync(m->dK.p, cur_kps, n * sizeof(pgb_keypoint), cudaMemcpyHostToDevice, s));
  PGB_CUDA(cudaMemcpyAsync(m->dD.p, cur_desc, n * 32, cudaMemcpyHostToDevice, s));
  PGB_CUDA(cudaMemcpyAsync(m->dQD.p, q_desc, n * 32, cudaMemcpyHostToDevice, s));
  PGB_CUDA(cudaMemcpyAsync(m->dQV.p, q_valid, n, cudaMemcpyHostToDevice, s));
  PGB_CUDA(cudaMemcpyAsync(m->dUV.p, q_uv, n * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
  PGB_CUDA(cudaMemcpyAsync(m->dAng.p, q_angle, n * sizeof(float), cudaMemcpyHostToDevice, s));
  PGB_CUDA(cudaMemcpyAsync(m->dOct.p, q_octave, n * sizeof(int), cudaMemcpyHostToDevice, s));
  PGB_CUDA(cudaMemcpyAsync(m->dN.p, cur_counts, n_pairs * sizeof(int), cudaMemcpyHostToDevice, s));
  PGB_CUDA(cudaMemcpyAsync(m->dQN.p, q_counts, n_pairs * sizeof(int), cudaMemcpyHostToDevice, s));
  A.curK = m->dK.p; A.curD = m->dD.p; A.curN = m->dN.p; A.qUV = m->dUV.p; A.qOct = m->dOct.p; A.qAng = m->dAng.p;
  A.qD = m->dQD.p; A.qValid = m->dQV.p; A.qN = m->dQN.p; A.matchOfCur = m->dMatch.p; A.nMatches = m->dNm.p;
  int rc = launch_match(m, A, n_pairs);
  if (rc) return rc;
  PGB_CUDA(cudaMemcpyAsync(match_of_cur, m->dMatch.p, n * sizeof(int), cudaMemcpyDeviceToHost, s));
  PGB_CUDA(cudaMemcpyAsync(n_matches, m->dNm.p, n_pairs * sizeof(int), cudaMemcpyDeviceToHost, s));
  PGB_CUDA(cudaStreamSynchronize(s));
  return PGB_OK;
}

int pgb_match_consecutive(pgb_matcher* m, int n_pairs, int cap, const pgb_keypoint* kps, const uint8_t* desc,
                          const int32_t* counts, const float* flow, float max_x, float max_y, float th,
                          const float* scale_factors, int nlevels, int32_t* match_of_cur, int32_t* n_matches) {
  if (!m) return fail(PGB_ERR_INVALID, "null handle");
  if (n_pairs < 0 || cap <= 0 || nlevels <= 0 || nlevels > 16 || !scale_factors || !(max_x > 0) || !(max_y > 0))
    return fail(PGB_ERR_INVALID, "pgb_match_consecutive: invalid argument");
  if (n_pairs == 0) return PGB_OK;
  if (!kps || !desc || !counts || !flow || !match_of_cur || !n_matches)
    return fail(PGB_ERR_INVALID, "pgb_match_consecutive: null buffer");
  PGB_CUDA(cudaSetDevice(m->device));
  MatchArgs A;
  fill_common(A, cap, 0.f, max_x, 0.f, max_y, th, scale_factors, nlevels, m->checkOri);
  A.consecutive = 1;
  A.curK = kps; A.curD = desc; A.curN = counts; A.flow = flow;
  A.matchOfCur = match_of_cur; A.nMatches = n_matches;
  int rc = launch_match(m, A, n_pairs);
  if (rc) return rc;
  // Tracking.cc:879-883: wider window when fewer than 20 matches; pairs that already have >= 20 exit at once.
  A.th = 2 * th;
  A.onlyIfBelow20 = 1;
  return launch_match(m, A, n_pairs);
}

int pgb_match_median_flow(int device, int n_pairs, int cap, const pgb_keypoint* kps, const int32_t* counts,
                          const int32_t* match_of_cur, const int32_t* n_matches, float* flow_xy, int32_t* tracked, void* stream) {
  if (n_pairs < 0 || cap <= 0 || cap > kFlowCap) return fail(PGB_ERR_INVALID, "pgb_match_median_flow: invalid argument (cap <= %d)", kFlowCap);
  if (n_pairs == 0) return PGB_OK;
  if (!kps || !counts || !match_of_cur || !n_matches || !flow_xy || !tracked)
    return fail(PGB_ERR_INVALID, "pgb_match_median_flow: null buffer");
  PGB_CUDA(cudaSetDevice(device));
  k_median_flow<<<n_pairs, 256, 0, (cudaStream_t)stream>>>(cap, kps, counts, match_of_cur, n_matches, flow_xy, tracked);
  PGB_CHECK_LAUNCH();
  return PGB_OK;
}

int pgb_descriptor_distance(const uint8_t* a, const uint8_t* b, int n, int32_t* dist, int is_device, void* stream) {
  if (n < 0 || (n > 0 && (!a || !b || !dist))) return fail(PGB_ERR_INVALID, "pgb_descriptor_distance: invalid argument");
  if (n == 0) return PGB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (is_device) {
    k_desc_distance<<<(n + 255) / 256, 256, 0, s>>>(a, b, n, dist);
    PGB_CHECK_LAUNCH();
    return PGB_OK;
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || use_device(dev)) return PGB_ERR_CUDA;
  DevBuf<uint8_t> da, db;
  DevBuf<int> dd;
  if (da.alloc((size_t)n * 32) || db.alloc((size_t)n * 32) || dd.alloc(n)) return PGB_ERR_CUDA;
  PGB_CUDA(cudaMemcpyAsync(da.p, a, (size_t)n * 32, cudaMemcpyHostToDevice, s));
  PGB_CUDA(cudaMemcpyAsync(db.p, b, (size_t)n * 32, cudaMemcpyHostToDevice, s));
  k_desc_distance<<<(n + 255) / 256, 256, 0, s>>>(da.p, db.p, n, dd.p);
  PGB_CHECK_LAUNCH();
  PGB_CUDA(cudaMemcpyAsync(dist, dd.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s));
  PGB_CUDA(cudaStreamSynchronize(s));
  return PGB_OK;
}

}  // extern "C"

namespace {

// host->device staging of one input array (or pass-through when the caller's buffers are device-resident)
template <typename B, typename T>
int stage_in(B& d, const T*& ptr, size_t n, bool is_device, cudaStream_t s) {
  if (is_device || !ptr) return PGB_OK;
  if (d.alloc(n)) return PGB_ERR_CUDA;
  PGB_CUDA(cudaMemcpyAsync(d.p, ptr, n * sizeof(T), cudaMemcpyHostToDevice, s));
  ptr = d.p;
  return PGB_OK;
}

int run_windowed(pgb_matcher* m, WinArgs& A, int nProb) {
  TempBuf<uint32_t> rows;
  TempBuf<int> cnt, err;
  const size_t n = (size_t)nProb * A.cap;
  if (rows.alloc(n * kWinK) || cnt.alloc(n) || err.alloc(1)) return PGB_ERR_CUDA;
  PGB_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), m->stream));
  A.rows = rows.p; A.rowCnt = cnt.p; A.err = err.p;
  const size_t smem = (size_t)A.cap * 16 + 64;
  if (smem > 227 * 1024) return fail(PGB_ERR_CAPACITY, "cap %d needs %zu B of shared memory (max 227 KB)", A.cap, smem);
  static DynSmemLimit limWindowed;
  if (int rc = limWindowed.ensure(k_match_windowed, smem)) return rc;
  k_match_windowed<<<nProb, kMtThreads, smem, m->stream>>>(A);
  PGB_CHECK_LAUNCH();
  int e = 0;
  PGB_CUDA(cudaMemcpyAsync(&e, err.p, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  PGB_CUDA(cudaStreamSynchronize(m->stream));
  if (e) return fail(PGB_ERR_CAPACITY, "a search window holds more than %d candidates", kWinK);
  return PGB_OK;
}

}  // namespace

extern "C" {

int pgb_match_for_initialization(pgb_matcher* m, int n_pairs, int cap, const pgb_keypoint* kps1, const uint8_t* desc1,
                                 const int32_t* counts1, const pgb_keypoint* kps2, const uint8_t* desc2,
                                 const int32_t* counts2, float* prev_matched_xy, int window_size, float min_x, float max_x,
                                 float min_y, float max_y, int32_t* matches12, int32_t* n_matches, int is_device) {
  if (!m) return fail(PGB_ERR_INVALID, "null handle");
  if (n_pairs < 0 || cap <= 0 || cap > 65535 || window_size <= 0 || !(max_x > min_x) || !(max_y > min_y))
    return fail(PGB_ERR_INVALID, "pgb_match_for_initialization: invalid argument");
  if (n_pairs == 0) return PGB_OK;
  if (!kps1 || !desc1 || !counts1 || !kps2 || !desc2 || !counts2 || !prev_matched_xy || !matches12 || !n_matches)
    return fail(PGB_ERR_INVALID, "pgb_match_for_initialization: null buffer");
  PGB_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->stream;
  TempScope scope(m->device, s);
  if (scope.rc) return PGB_ERR_CUDA;
  const size_t n = (size_t)n_pairs * cap;
  // the kernel wants the F1 keypoints as separate level / angle arrays
  std::vector<pgb_keypoint> hk1;
  TempBuf<pgb_keypoint> dk1tmp;
  const pgb_keypoint* k1host = kps1;
  if (is_device) {
    hk1.resize(n);
    PGB_CUDA(cudaMemcpyAsync(hk1.data(), kps1, n * sizeof(pgb_keypoint), cudaMemcpyDeviceToHost, s));
    PGB_CUDA(cudaStreamSynchronize(s));
    k1host = hk1.data();
  }
  std::vector<int> lvl(n);
  std::vector<float> ang(n);
  for (size_t i = 0; i < n; i++) { lvl[i] = k1host[i].octave; ang[i] = k1host[i].angle; }
  TempBuf<int> dLvl, dN1, dN2, dM12, dNm;
  TempBuf<float> dAng, dUV, dUVOut;
  TempBuf<pgb_keypoint> dK2;
  TempBuf<uint8_t> dD1, dD2;
  if (dLvl.alloc(n) || dAng.alloc(n) || dUVOut.alloc(2 * n)) return PGB_ERR_CUDA;
  PGB_CUDA(cudaMemcpyAsync(dLvl.p, lvl.data(), n * sizeof(int), cudaMemcpyHostToDevice, s));
  PGB_CUDA(cudaMemcpyAsync(dAng.p, ang.data(), n * sizeof(float), cudaMemcpyHostToDevice, s));
  const float* uv = prev_matched_xy;
  int rc = stage_in(dK2, kps2, n, is_device, s) | stage_in(dD1, desc1, n * 32, is_device, s) | stage_in(dD2, desc2, n * 32, is_device, s) |
           stage_in(dN1, counts1, n_pairs, is_device, s) | stage_in(dN2, counts2, n_pairs, is_device, s) | stage_in(dUV, uv, 2 * n, is_device, s);
  if (rc) return PGB_ERR_CUDA;
  int* dm12 = matches12;
  int* dnm = n_matches;
  if (!is_device) {
    if (dM12.alloc(n) || dNm.alloc(n_pairs)) return PGB_ERR_CUDA;
    dm12 = dM12.p; dnm = dNm.p;
  }
  WinArgs A;
  memset(&A, 0, sizeof A);
  A.flavour = 1; A.cap = cap; A.nlevels = 1; A.checkOri = m->checkOri; A.nnratio = m->nnratio;
  A.minX = min_x; A.maxX = max_x; A.minY = min_y; A.maxY = max_y; A.th = (float)window_size;
  A.curK = kps2; A.curD = desc2; A.curN = counts2; A.qUV = uv; A.qLevel = dLvl.p; A.qAux = dAng.p; A.qD = desc1; A.qN = counts1;
  A.matchOfQ = dm12; A.qUVOut = dUVOut.p; A.nMatches = dnm;
  rc = run_windowed(m, A, n_pairs);
  if (rc) return rc;
  PGB_CUDA(cudaMemcpyAsync(prev_matched_xy, dUVOut.p, 2 * n * sizeof(float), is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  if (!is_device) {
    PGB_CUDA(cudaMemcpyAsync(matches12, dM12.p, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    PGB_CUDA(cudaMemcpyAsync(n_matches, dNm.p, n_pairs * sizeof(int), cudaMemcpyDeviceToHost, s));
  }
  PGB_CUDA(cudaStreamSynchronize(s));
  return PGB_OK;
}

int pgb_match_map_points(pgb_matcher* m, int n_frames, int cap, const pgb_keypoint* kps, const uint8_t* desc,
                         const int32_t* counts, const uint8_t* has_map_point, const float* proj_xy,
                         const int32_t* track_level, const float* view_cos, const uint8_t* mp_desc, const uint8_t* in_view,
                         const uint8_t* mp_observed, const int32_t* mp_counts, float min_x, float max_x, float min_y,
                         float max_y, float th, const float* scale_factors, int nlevels, int32_t* match_of_feature,
                         int32_t* n_matches, int is_device) {
  if (!m) return fail(PGB_ERR_INVALID, "null handle");
  if (n_frames < 0 || cap <= 0 || cap > 65535 || nlevels <= 0 || nlevels > 16 || !scale_factors || !(max_x > min_x) || !(max_y > min_y))
    return fail(PGB_ERR_INVALID, "pgb_match_map_points: invalid argument");
  if (n_frames == 0) return PGB_OK;
  if (!kps || !desc || !counts || !proj_xy || !track_level || !view_cos || !mp_desc || !in_view || !mp_observed || !mp_counts ||
      !match_of_feature || !n_matches)
    return fail(PGB_ERR_INVALID, "pgb_match_map_points: null buffer");
  PGB_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->stream;
  TempScope scope(m->device, s);
  if (scope.rc) return PGB_ERR_CUDA;
  const size_t n = (size_t)n_frames * cap;
  TempBuf<pgb_keypoint> dK;
  TempBuf<uint8_t> dD, dHas, dQD, dView, dObs;
  TempBuf<float> dUV, dCos;
  TempBuf<int> dN, dLvl, dQN, dMatch, dNm;
  int rc = stage_in(dK, kps, n, is_device, s) | stage_in(dD, desc, n * 32, is_device, s) | stage_in(dN, counts, n_frames, is_device, s) |
           stage_in(dHas, has_map_point, n, is_device, s) | stage_in(dUV, proj_xy, 2 * n, is_device, s) |
           stage_in(dLvl, track_level, n, is_device, s) | stage_in(dCos, view_cos, n, is_device, s) |
           stage_in(dQD, mp_desc, n * 32, is_device, s) | stage_in(dView, in_view, n, is_device, s) |
           stage_in(dObs, mp_observed, n, is_device, s) | stage_in(dQN, mp_counts, n_frames, is_device, s);
  if (rc) return PGB_ERR_CUDA;
  int* dmatch = match_of_feature;
  int* dnm = n_matches;
  if (!is_device) {
    if (dMatch.alloc(n) || dNm.alloc(n_frames)) return PGB_ERR_CUDA;
    dmatch = dMatch.p; dnm = dNm.p;
  }
  WinArgs A;
  memset(&A, 0, sizeof A);
  A.flavour = 2; A.cap = cap; A.nlevels = nlevels; A.checkOri = 0; A.nnratio = m->nnratio;
  A.minX = min_x; A.maxX = max_x; A.minY = min_y; A.maxY = max_y; A.th = th;
  for (int i = 0; i < nlevels; i++) A.scale[i] = scale_factors[i];
  A.curK = kps; A.curD = desc; A.curN = counts; A.curTaken = has_map_point; A.qUV = proj_xy; A.qLevel = track_level;
  A.qAux = view_cos; A.qD = mp_desc; A.qValid = in_view; A.qObs = mp_observed; A.qN = mp_counts;
  A.matchOfCur = dmatch; A.nMatches = dnm;
  rc = run_windowed(m, A, n_frames);
  if (rc) return rc;
  if (!is_device) {
    PGB_CUDA(cudaMemcpyAsync(match_of_feature, dMatch.p, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    PGB_CUDA(cudaMemcpyAsync(n_matches, dNm.p, n_frames * sizeof(int), cudaMemcpyDeviceToHost, s));
    PGB_CUDA(cudaStreamSynchronize(s));
  }
  return PGB_OK;
}

int pgb_match_by_bow(pgb_matcher* m, int n_pairs, int cap, const uint8_t* kf_desc, const float* kf_angle,
                     const uint8_t* kf_has_map_point, const int32_t* kf_node_off, const uint32_t* kf_node_id,
                     const int32_t* kf_feat_start, const uint32_t* kf_feat_idx, int kf_nodes_total, int kf_idx_total,
                     const uint8_t* f_desc, const float* f_angle, const int32_t* f_counts, const int32_t* f_node_off,
                     const uint32_t* f_node_id, const int32_t* f_feat_start, const uint32_t* f_feat_idx, int f_nodes_total,
                     int f_idx_total, int32_t* match_of_feature, int32_t* n_matches, int is_device) {
  if (!m) return fail(PGB_ERR_INVALID, "null handle");
  if (n_pairs < 0 || cap <= 0 || cap > 65535 || kf_nodes_total < 0 || kf_idx_total < 0 || f_nodes_total < 0 || f_idx_total < 0)
    return fail(PGB_ERR_INVALID, "pgb_match_by_bow: invalid argument");
  if (n_pairs == 0) return PGB_OK;
  if (!kf_desc || !kf_angle || !kf_has_map_point || !kf_node_off || !kf_feat_start || !f_desc || !f_angle || !f_counts ||
      !f_node_off || !f_feat_start || !match_of_feature || !n_matches || (kf_nodes_total && !kf_node_id) ||
      (kf_idx_total && !kf_feat_idx) || (f_nodes_total && !f_node_id) || (f_idx_total && !f_feat_idx))
    return fail(PGB_ERR_INVALID, "pgb_match_by_bow: null buffer");
  PGB_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->stream;
  TempScope scope(m->device, s);
  if (scope.rc) return PGB_ERR_CUDA;
  const size_t n = (size_t)n_pairs * cap;
  TempBuf<uint8_t> dKD, dKH, dFD;
  TempBuf<float> dKA, dFA;
  TempBuf<int> dKO, dKS, dFN, dFO, dFS, dMatch, dNm, dErr;
  TempBuf<uint32_t> dKN, dKI, dFNode, dFI;
  int rc = stage_in(dKD, kf_desc, n * 32, is_device, s) | stage_in(dKA, kf_angle, n, is_device, s) |
           stage_in(dKH, kf_has_map_point, n, is_device, s) | stage_in(dKO, kf_node_off, (size_t)n_pairs + 1, is_device, s) |
           stage_in(dKN, kf_node_id, kf_nodes_total, is_device, s) | stage_in(dKS, kf_feat_start, (size_t)kf_nodes_total + 1, is_device, s) |
           stage_in(dKI, kf_feat_idx, kf_idx_total, is_device, s) | stage_in(dFD, f_desc, n * 32, is_device, s) |
           stage_in(dFA, f_angle, n, is_device, s) | stage_in(dFN, f_counts, n_pairs, is_device, s) |
           stage_in(dFO, f_node_off, (size_t)n_pairs + 1, is_device, s) | stage_in(dFNode, f_node_id, f_nodes_total, is_device, s) |
           stage_in(dFS, f_feat_start, (size_t)f_nodes_total + 1, is_device, s) | stage_in(dFI, f_feat_idx, f_idx_total, is_device, s);
  if (rc || dErr.alloc(1)) return PGB_ERR_CUDA;
  PGB_CUDA(cudaMemsetAsync(dErr.p, 0, sizeof(int), s));
  int* dmatch = match_of_feature;
  int* dnm = n_matches;
  if (!is_device) {
    if (dMatch.alloc(n) || dNm.alloc(n_pairs)) return PGB_ERR_CUDA;
    dmatch = dMatch.p; dnm = dNm.p;
  }
  BowArgs A;
  memset(&A, 0, sizeof A);
  A.cap = cap; A.checkOri = m->checkOri; A.nnratio = m->nnratio;
  A.kfD = kf_desc; A.kfAng = kf_angle; A.kfHas = kf_has_map_point; A.kfNodeOff = kf_node_off; A.kfNode = kf_node_id;
  A.kfStart = kf_feat_start; A.kfIdx = kf_feat_idx; A.fD = f_desc; A.fAng = f_angle; A.fN = f_counts; A.fNodeOff = f_node_off;
  A.fNode = f_node_id; A.fStart = f_feat_start; A.fIdx = f_feat_idx; A.matchOfF = dmatch; A.nMatches = dnm; A.err = dErr.p;
  const size_t smem = 128 + (size_t)((cap + 31) / 32) * 4 + cap + 16;
  if (smem > 227 * 1024) return fail(PGB_ERR_CAPACITY, "cap %d needs %zu B of shared memory (max 227 KB)", cap, smem);
  static DynSmemLimit limBow;
  if (int rc = limBow.ensure(k_match_bow, smem)) return rc;
  k_match_bow<<<n_pairs, kMtThreads, smem, s>>>(A);
  PGB_CHECK_LAUNCH();
  int e = 0;
  PGB_CUDA(cudaMemcpyAsync(&e, dErr.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  if (!is_device) {
    PGB_CUDA(cudaMemcpyAsync(match_of_feature, dMatch.p, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    PGB_CUDA(cudaMemcpyAsync(n_matches, dNm.p, n_pairs * sizeof(int), cudaMemcpyDeviceToHost, s));
  }
  PGB_CUDA(cudaStreamSynchronize(s));
  if (e) return fail(PGB_ERR_INVALID, "pgb_match_by_bow: feature vectors must have strictly increasing node ids, in-range indices and "
                                      "every frame feature in at most one node");
  return PGB_OK;
}

int pgb_distinctive_descriptors(const uint8_t* desc, const int32_t* offsets, int n_points, int32_t* best_idx, int is_device,
                                void* stream) {
  if (n_points < 0 || (n_points > 0 && (!desc || !offsets || !best_idx)))
    return fail(PGB_ERR_INVALID, "pgb_distinctive_descriptors: invalid argument");
  if (n_points == 0) return PGB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || use_device(dev)) return PGB_ERR_CUDA;
  TempScope scope(dev, s);
  if (scope.rc) return PGB_ERR_CUDA;
  TempBuf<uint8_t> dD;
  TempBuf<int> dO, dB, dE;
  if (dE.alloc(1)) return PGB_ERR_CUDA;
  PGB_CUDA(cudaMemsetAsync(dE.p, 0, sizeof(int), s));
  const uint8_t* pd = desc;
  const int* po = offsets;
  int* pb = best_idx;
  if (!is_device) {
    const int total = offsets[n_points];
    if (total < 0) return fail(PGB_ERR_INVALID, "offsets must be non-decreasing");
    if (dD.alloc((size_t)std::max(total, 1) * 32) || dO.alloc(n_points + 1) || dB.alloc(n_points)) return PGB_ERR_CUDA;
    PGB_CUDA(cudaMemcpyAsync(dD.p, desc, (size_t)total * 32, cudaMemcpyHostToDevice, s));
    PGB_CUDA(cudaMemcpyAsync(dO.p, offsets, (n_points + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    pd = dD.p; po = dO.p; pb = dB.p;
  }
  k_distinctive<<<(n_points + 3) / 4, 128, 0, s>>>(pd, po, n_points, pb, dE.p);
  PGB_CHECK_LAUNCH();
  int e = 0;
  PGB_CUDA(cudaMemcpyAsync(&e, dE.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  if (!is_device) PGB_CUDA(cudaMemcpyAsync(best_idx, dB.p, n_points * sizeof(int), cudaMemcpyDeviceToHost, s));
  PGB_CUDA(cudaStreamSynchronize(s));
  if (e) return fail(PGB_ERR_CAPACITY, "a map point has more than %d observations", kDistinctMaxN);
  return PGB_OK;
}

}  // extern "C"
