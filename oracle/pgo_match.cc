// ORACLE -- TEST INFRASTRUCTURE ONLY (see pgo_orb.cc header for the rules).
// CPU restatement of the reference's frame-to-frame projection matcher:
//   thirdparty/orb-slam2/src/ORBmatcher.cc  SearchByProjection(Frame&,const Frame&,th,bMono) :1332-1474,
//   ComputeThreeMaxima :1605-1646, DescriptorDistance :1651-1667, constants :38-40
//   thirdparty/orb-slam2/src/Frame.cc       AssignFeaturesToGrid :234-249, GetFeaturesInArea :331-384,
//   PosInGrid :386-396; grid 64x48 (Frame.h:37-38)
//   thirdparty/orb-slam2/src/Tracking.cc    retry with 2*th when fewer than 20 matches :876-883
// The Frame/MapPoint object graph is flattened to arrays: every query is a "last frame map point" already
// projected to (u,v) with its octave, angle and representative descriptor.  Mono only (bForward/bBackward are
// false, mvuRight = -1).  The reference has no tests for this path.  PINNED against the reference's own source: `make -C
// oracle _ref` compiles the bodies of these functions from ORBmatcher.cc / Frame.cc (stand-in class declarations in
// ref_shims/pgo_orbslam_shim.h) and tests/test_oracle_reference_pin.py holds every flavour restated here identical to them
// (match vectors, counts, updated vbPrevMatched); further: the brute-force property tests in tests/test_oracle_match.py.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "pgo.h"

namespace {
const int TH_HIGH = 100;
const int HISTO_LENGTH = 30;
const int GRID_ROWS = 48, GRID_COLS = 64;

int descriptor_distance(const uint8_t* a, const uint8_t* b) {
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t pa, pb;
    memcpy(&pa, a + 4 * i, 4);
    memcpy(&pb, b + 4 * i, 4);
    uint32_t v = pa ^ pb;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = (int)histo[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}
}  // namespace

extern "C" {

int pgo_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

int pgo_search_by_projection(const pgb_keypoint* cur_kps, const uint8_t* cur_desc, int n_cur, const float* q_uv,
                             const int32_t* q_octave, const float* q_angle, const uint8_t* q_desc,
                             const uint8_t* q_valid, int n_q, float minX, float maxX, float minY, float maxY, float th,
                             const float* scale_factors, int nlevels, int check_ori, int32_t* match_of_cur,
                             int32_t* best_dist_of_q) {
  (void)nlevels;
  const float invW = (float)GRID_COLS / (maxX - minX);
  const float invH = (float)GRID_ROWS / (maxY - minY);
  // AssignFeaturesToGrid
  static thread_local std::vector<int> grid[GRID_COLS][GRID_ROWS];
  for (int i = 0; i < GRID_COLS; i++)
    for (int j = 0; j < GRID_ROWS; j++) grid[i][j].clear();
  for (int i = 0; i < n_cur; i++) {
    int posX = (int)std::round((cur_kps[i].x - minX) * invW);
    int posY = (int)std::round((cur_kps[i].y - minY) * invH);
    if (posX < 0 || posX >= GRID_COLS || posY < 0 || posY >= GRID_ROWS) continue;
    grid[posX][posY].push_back(i);
  }
  for (int i = 0; i < n_cur; i++) match_of_cur[i] = -1;
  if (best_dist_of_q)
    for (int i = 0; i < n_q; i++) best_dist_of_q[i] = -1;

  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  std::vector<int> vIndices2;
  for (int i = 0; i < n_q; i++) {
    if (!q_valid[i]) continue;
    const float u = q_uv[2 * i], v = q_uv[2 * i + 1];
    if (u < minX || u > maxX) continue;
    if (v < minY || v > maxY) continue;
    const int nLastOctave = q_octave[i];
    const float r = th * scale_factors[nLastOctave];
    const int minLevel = nLastOctave - 1, maxLevel = nLastOctave + 1;
    // GetFeaturesInArea
    vIndices2.clear();
    do {
      const int nMinCellX = std::max(0, (int)std::floor((u - minX - r) * invW));
      if (nMinCellX >= GRID_COLS) break;
      const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((u - minX + r) * invW));
      if (nMaxCellX < 0) break;
      const int nMinCellY = std::max(0, (int)std::floor((v - minY - r) * invH));
      if (nMinCellY >= GRID_ROWS) break;
      const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((v - minY + r) * invH));
      if (nMaxCellY < 0) break;
      const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
      for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
          for (int idx : grid[ix][iy]) {
            const pgb_keypoint& kp = cur_kps[idx];
            if (bCheckLevels) {
              if (kp.octave < minLevel) continue;
              if (maxLevel >= 0 && kp.octave > maxLevel) continue;
            }
            const float distx = kp.x - u, disty = kp.y - v;
            if (std::fabs(distx) < r && std::fabs(disty) < r) vIndices2.push_back(idx);
          }
    } while (0);
    if (vIndices2.empty()) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (int i2 : vIndices2) {
      if (match_of_cur[i2] >= 0) continue;  // already holds a map point with observations
      const int dist = descriptor_distance(q_desc + (size_t)i * 32, cur_desc + (size_t)i2 * 32);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (best_dist_of_q) best_dist_of_q[i] = bestDist;
    if (bestDist <= TH_HIGH) {
      match_of_cur[bestIdx2] = i;
      nmatches++;
      if (check_ori) {
        float rot = q_angle[i] - cur_kps[bestIdx2].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)std::round(rot * factor);
        if (bin == HISTO_LENGTH) bin = 0;
        rotHist[bin].push_back(bestIdx2);
      }
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++)
      if (i != ind1 && i != ind2 && i != ind3)
        for (int idx : rotHist[i]) { match_of_cur[idx] = -1; nmatches--; }
  }
  return nmatches;
}

// The synthetic benchmark's match stage: previous frame's keypoints are the map points, projected
// by the known flow; retry at 2*th when fewer than 20 matches (Tracking.cc:876-883).
namespace {
// Frame::AssignFeaturesToGrid + GetFeaturesInArea (Frame.cc:234-249, 331-384), shared by the flavours below.
struct FrameGrid {
  std::vector<int> cell[GRID_COLS][GRID_ROWS];
  float minX, minY, invW, invH;
  const pgb_keypoint* kps;
  FrameGrid(const pgb_keypoint* k, int n, float mnX, float mxX, float mnY, float mxY) : minX(mnX), minY(mnY), kps(k) {
    invW = (float)GRID_COLS / (mxX - mnX);
    invH = (float)GRID_ROWS / (mxY - mnY);
    for (int i = 0; i < n; i++) {
      const int posX = (int)std::round((k[i].x - minX) * invW), posY = (int)std::round((k[i].y - minY) * invH);
      if (posX < 0 || posX >= GRID_COLS || posY < 0 || posY >= GRID_ROWS) continue;
      cell[posX][posY].push_back(i);
    }
  }
  std::vector<size_t> in_area(float x, float y, float r, int minLevel, int maxLevel) const {
    std::vector<size_t> out;
    const int nMinCellX = std::max(0, (int)std::floor((x - minX - r) * invW));
    if (nMinCellX >= GRID_COLS) return out;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - minX + r) * invW));
    if (nMaxCellX < 0) return out;
    const int nMinCellY = std::max(0, (int)std::floor((y - minY - r) * invH));
    if (nMinCellY >= GRID_ROWS) return out;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - minY + r) * invH));
    if (nMaxCellY < 0) return out;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
      for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
        for (int idx : cell[ix][iy]) {
          const pgb_keypoint& kp = kps[idx];
          if (bCheckLevels) {
            if (kp.octave < minLevel) continue;
            if (maxLevel >= 0)
              if (kp.octave > maxLevel) continue;
          }
          const float distx = kp.x - x, disty = kp.y - y;
          if (std::fabs(distx) < r && std::fabs(disty) < r) out.push_back(idx);
        }
    return out;
  }
};
const int TH_LOW = 50;
}  // namespace

// ORBmatcher::SearchForInitialization (ORBmatcher.cc:407-522), literal.  prev_matched is updated in place.
int pgo_search_for_initialization(const pgb_keypoint* k1, const uint8_t* d1, int n1, const pgb_keypoint* k2,
                                  const uint8_t* d2, int n2, float* prev_matched, int windowSize, float minX, float maxX,
                                  float minY, float maxY, float nnratio, int check_ori, int32_t* vnMatches12) {
  int nmatches = 0;
  for (int i = 0; i < n1; i++) vnMatches12[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  std::vector<int> vMatchedDistance(n2, INT32_MAX), vnMatches21(n2, -1);
  const FrameGrid F2(k2, n2, minX, maxX, minY, maxY);
  for (int i1 = 0; i1 < n1; i1++) {
    const int level1 = k1[i1].octave;
    if (level1 > 0) continue;
    const std::vector<size_t> vIndices2 = F2.in_area(prev_matched[2 * i1], prev_matched[2 * i1 + 1], (float)windowSize, level1, level1);
    if (vIndices2.empty()) continue;
    int bestDist = INT32_MAX, bestDist2 = INT32_MAX, bestIdx2 = -1;
    for (size_t i2 : vIndices2) {
      const int dist = descriptor_distance(d1 + (size_t)i1 * 32, d2 + i2 * 32);
      if (vMatchedDistance[i2] <= dist) continue;
      if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = (int)i2; }
      else if (dist < bestDist2) { bestDist2 = dist; }
    }
    if (bestDist <= TH_LOW) {
      if (bestDist < (float)bestDist2 * nnratio) {
        if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
        vnMatches12[i1] = bestIdx2;
        vnMatches21[bestIdx2] = i1;
        vMatchedDistance[bestIdx2] = bestDist;
        nmatches++;
        if (check_ori) {
          float rot = k1[i1].angle - k2[bestIdx2].angle;
          if (rot < 0.0) rot += 360.0f;
          int bin = (int)std::round(rot * factor);
          if (bin == HISTO_LENGTH) bin = 0;
          rotHist[bin].push_back(i1);
        }
      }
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i])
        if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
    }
  }
  for (int i1 = 0; i1 < n1; i1++)
    if (vnMatches12[i1] >= 0) { prev_matched[2 * i1] = k2[vnMatches12[i1]].x; prev_matched[2 * i1 + 1] = k2[vnMatches12[i1]].y; }
  return nmatches;
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (ORBmatcher.cc:46-131), literal, mono.
// match_of_feature[idx] = index of the map point this call put into F.mvpMapPoints[idx], else -1.
int pgo_search_map_points(const pgb_keypoint* kps, const uint8_t* desc, int n, const uint8_t* has_map_point,
                          const float* proj_xy, const int32_t* track_level, const float* view_cos, const uint8_t* mp_desc,
                          const uint8_t* in_view, const uint8_t* mp_observed, int n_mp, float minX, float maxX, float minY,
                          float maxY, float th, const float* scale_factors, float nnratio, int32_t* match_of_feature) {
  int nmatches = 0;
  const bool bFactor = th != 1.0;
  const FrameGrid F(kps, n, minX, maxX, minY, maxY);
  std::vector<uint8_t> observed(has_map_point, has_map_point + n);  // F.mvpMapPoints[idx] && Observations() > 0
  for (int i = 0; i < n; i++) match_of_feature[i] = -1;
  for (int iMP = 0; iMP < n_mp; iMP++) {
    if (!in_view[iMP]) continue;  // !mbTrackInView || isBad()
    const int nPredictedLevel = track_level[iMP];
    float r = view_cos[iMP] > 0.998 ? 2.5f : 4.0f;
    if (bFactor) r *= th;
    const std::vector<size_t> vIndices = F.in_area(proj_xy[2 * iMP], proj_xy[2 * iMP + 1], r * scale_factors[nPredictedLevel],
                                                   nPredictedLevel - 1, nPredictedLevel);
    if (vIndices.empty()) continue;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (size_t idx : vIndices) {
      if (observed[idx]) continue;
      const int dist = descriptor_distance(mp_desc + (size_t)iMP * 32, desc + idx * 32);
      if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kps[idx].octave; bestIdx = (int)idx; }
      else if (dist < bestDist2) { bestLevel2 = kps[idx].octave; bestDist2 = dist; }
    }
    if (bestDist <= TH_HIGH) {
      if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
      match_of_feature[bestIdx] = iMP;
      observed[bestIdx] = mp_observed[iMP];
      nmatches++;
    }
  }
  return nmatches;
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (ORBmatcher.cc:161-290), literal.  The two
// DBoW2::FeatureVector maps (node id -> feature indices, std::map order) arrive as CSR arrays sorted by node id; the
// merge with lower_bound jumps is restated on them.  kf_has_map_point[i] = vpMapPointsKF[i] && !isBad().
// match_of_feature[idxF] = index of the keyframe feature whose map point F's feature received, else -1.
int pgo_search_by_bow(const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_has_map_point,
                      const uint32_t* kf_node_id, const int32_t* kf_feat_start, const uint32_t* kf_feat_idx, int kf_nodes,
                      const uint8_t* f_desc, const float* f_angle, int f_n, const uint32_t* f_node_id,
                      const int32_t* f_feat_start, const uint32_t* f_feat_idx, int f_nodes, float nnratio, int check_ori,
                      int32_t* match_of_feature) {
  for (int i = 0; i < f_n; i++) match_of_feature[i] = -1;
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  int KFit = 0, Fit = 0;
  while (KFit != kf_nodes && Fit != f_nodes) {
    if (kf_node_id[KFit] == f_node_id[Fit]) {
      for (int iKF = kf_feat_start[KFit]; iKF < kf_feat_start[KFit + 1]; iKF++) {
        const unsigned realIdxKF = kf_feat_idx[iKF];
        if (!kf_has_map_point[realIdxKF]) continue;
        const uint8_t* dKF = kf_desc + (size_t)realIdxKF * 32;
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        for (int iF = f_feat_start[Fit]; iF < f_feat_start[Fit + 1]; iF++) {
          const unsigned realIdxF = f_feat_idx[iF];
          if (match_of_feature[realIdxF] >= 0) continue;
          const int dist = descriptor_distance(dKF, f_desc + (size_t)realIdxF * 32);
          if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = (int)realIdxF; }
          else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 <= TH_LOW) {
          if ((float)bestDist1 < nnratio * (float)bestDist2) {
            match_of_feature[bestIdxF] = (int)realIdxKF;
            if (check_ori) {
              float rot = kf_angle[realIdxKF] - f_angle[bestIdxF];
              if (rot < 0.0) rot += 360.0f;
              int bin = (int)roundf(rot * factor);
              if (bin == HISTO_LENGTH) bin = 0;
              rotHist[bin].push_back(bestIdxF);
            }
            nmatches++;
          }
        }
      }
      KFit++; Fit++;
    } else if (kf_node_id[KFit] < f_node_id[Fit]) {
      KFit = (int)(std::lower_bound(kf_node_id, kf_node_id + kf_nodes, f_node_id[Fit]) - kf_node_id);
    } else {
      Fit = (int)(std::lower_bound(f_node_id, f_node_id + f_nodes, kf_node_id[KFit]) - f_node_id);
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx : rotHist[i]) { match_of_feature[idx] = -1; nmatches--; }
    }
  }
  return nmatches;
}

// MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:259-324), literal: index of the descriptor with the least
// median distance to the rest (float distance matrix, std::sort of each row, vDists[0.5*(N-1)]).
int pgo_distinctive_descriptor(const uint8_t* desc, int N) {
  if (N <= 0) return -1;
  std::vector<std::vector<float>> Distances(N, std::vector<float>(N));
  for (int i = 0; i < N; i++) {
    Distances[i][i] = 0;
    for (int j = i + 1; j < N; j++) {
      const int distij = descriptor_distance(desc + (size_t)i * 32, desc + (size_t)j * 32);
      Distances[i][j] = (float)distij;
      Distances[j][i] = (float)distij;
    }
  }
  int BestMedian = INT32_MAX, BestIdx = 0;
  for (int i = 0; i < N; i++) {
    std::vector<int> vDists(Distances[i].begin(), Distances[i].end());
    std::sort(vDists.begin(), vDists.end());
    const int median = vDists[(size_t)(0.5 * (N - 1))];
    if (median < BestMedian) { BestMedian = median; BestIdx = i; }
  }
  return BestIdx;
}

int pgo_match_consecutive(const pgb_keypoint* prev_kps, const uint8_t* prev_desc, int n_prev,
                          const pgb_keypoint* cur_kps, const uint8_t* cur_desc, int n_cur, float flow_x, float flow_y,
                          float maxX, float maxY, float th, const float* scale_factors, int nlevels,
                          int32_t* match_of_cur) {
  std::vector<float> uv(2 * (size_t)n_prev), ang(n_prev);
  std::vector<int32_t> oct(n_prev);
  std::vector<uint8_t> valid(n_prev, 1);
  for (int i = 0; i < n_prev; i++) {
    uv[2 * i] = prev_kps[i].x + flow_x;
    uv[2 * i + 1] = prev_kps[i].y + flow_y;
    ang[i] = prev_kps[i].angle;
    oct[i] = prev_kps[i].octave;
  }
  int n = pgo_search_by_projection(cur_kps, cur_desc, n_cur, uv.data(), oct.data(), ang.data(), prev_desc,
                                   valid.data(), n_prev, 0.f, maxX, 0.f, maxY, th, scale_factors, nlevels, 1,
                                   match_of_cur, nullptr);
  if (n < 20)
    n = pgo_search_by_projection(cur_kps, cur_desc, n_cur, uv.data(), oct.data(), ang.data(), prev_desc,
                                 valid.data(), n_prev, 0.f, maxX, 0.f, maxY, 2 * th, scale_factors, nlevels, 1,
                                 match_of_cur, nullptr);
  return n;
}

}  // extern "C"
