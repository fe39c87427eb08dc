// ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's ORB extraction path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
// The product (libpgb200.so) never links, loads or calls anything in oracle/.
//
// Follows (file:line relative to waiwnf/pilotguru):
//   thirdparty/orb-slam2/src/ORBextractor.cc  (ctor :410-470, IC_Angle :77-104, computeOrbDescriptor :108-147,
//   DivideNode :481-537, DistributeOctTree :539-763, ComputeKeyPointsOctTree :765-852, operator() :1042-1104,
//   ComputePyramid :1106-1131)
// and restates the un-vendored OpenCV primitives those lines call (cv::resize INTER_LINEAR, cv::FAST TYPE_9_16
// with NMS, cv::GaussianBlur 7x7 sigma 2, cv::fastAtan2, cvRound) from their published algorithms; these are
// pinned against cv2 4.13 (the only OpenCV in this image; the reference pins 2.4.x) by tests/test_oracle_orb.py
// and the fixtures under tests/golden/.  The reference ships no tests or golden vectors for this path
// (SURVEY.md section 4), so those pins are the anchor.
//
// Documented deviation: DistributeOctTree sorts (size, node pointer) pairs, i.e. breaks size ties by heap
// address (:684).  The oracle breaks ties by node creation sequence number (ascending sort, processed from the
// back => later-created first), SURVEY.md App. A.3.  That is what the reference itself does whenever its allocator
// hands out increasing addresses.
//
// PINNED against the reference's own source: `make -C oracle _ref` compiles thirdparty/orb-slam2/src/ORBextractor.cc
// where it lies (OpenCV stand-in in ref_shims/: the file's own logic runs from source, the OpenCV primitives it calls
// are this file's cv2-pinned restatements; monotonic allocator in ref_bump_alloc.cc), and
// tests/test_oracle_reference_pin.py holds this file's pyramids, all 7 keypoint fields and all descriptors identical
// to it on synthetic, noise, odd-sized and low-texture images.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <ctime>
#include <list>
#include <vector>

#include "../include/pgb200.h"
#include "../include/pgb200_orb_pattern.h"
#include "pgo.h"

namespace {

const int PATCH_SIZE = 31;
const int HALF_PATCH_SIZE = 15;
const int EDGE_THRESHOLD = 19;

inline int cv_round_f(float v) { return (int)lrintf(v); }   // cvRound: round-half-even
inline int cv_round_d(double v) { return (int)lrint(v); }

struct Image {
  int w = 0, h = 0;
  std::vector<uint8_t> d;
  Image() {}
  Image(int w_, int h_) : w(w_), h(h_), d((size_t)w_ * h_) {}
  uint8_t at(int y, int x) const { return d[(size_t)y * w + x]; }
};

// ---------------------------------------------------------------- cv::resize, INTER_LINEAR, 8UC1 (App. A.1)
void resize_linear(const uint8_t* src, int sw, int sh, size_t spitch, uint8_t* dst, int dw, int dh, size_t dpitch) {
  const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
  const double scale_x = 1.0 / inv_scale_x, scale_y = 1.0 / inv_scale_y;
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<short> ialpha(2 * dw), ibeta(2 * dh);
  for (int dx = 0; dx < dw; dx++) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)std::floor(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx;
    ialpha[2 * dx] = (short)cv_round_f((1.f - fx) * 2048);
    ialpha[2 * dx + 1] = (short)cv_round_f(fx * 2048);
  }
  for (int dy = 0; dy < dh; dy++) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)std::floor(fy);
    fy -= sy;
    // OpenCV clips the row index when fetching rows, not the coefficient (resizeGeneric_: clip(sy+k)).
    yofs[dy] = sy;
    ibeta[2 * dy] = (short)cv_round_f((1.f - fy) * 2048);
    ibeta[2 * dy + 1] = (short)cv_round_f(fy * 2048);
  }
  std::vector<int> row0(dw), row1(dw);
  for (int dy = 0; dy < dh; dy++) {
    int sy0 = std::min(std::max(yofs[dy], 0), sh - 1);
    int sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
    const uint8_t* S0 = src + (size_t)sy0 * spitch;
    const uint8_t* S1 = src + (size_t)sy1 * spitch;
    for (int dx = 0; dx < dw; dx++) {
      int sx = xofs[dx];
      int sx1 = std::min(sx + 1, sw - 1);
      int a0 = ialpha[2 * dx], a1 = ialpha[2 * dx + 1];
      row0[dx] = S0[sx] * a0 + S0[sx1] * a1;
      row1[dx] = S1[sx] * a0 + S1[sx1] * a1;
    }
    int b0 = ibeta[2 * dy], b1 = ibeta[2 * dy + 1];
    uint8_t* D = dst + (size_t)dy * dpitch;
    for (int dx = 0; dx < dw; dx++)
      D[dx] = (uint8_t)((((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2);
  }
}

// ---------------------------------------------------------------- cv::FAST TYPE_9_16 (App. A.2)
const int RING_DX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int RING_DY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// "bam": max over the 16 contiguous 9-arcs of max(min(v - p_k), min(p_k - v)).  Corner at th <=> bam > th;
// the score OpenCV reports (cornerScore<16>) is bam - 1.
int fast_bam(const uint8_t* p, const ptrdiff_t* ring) {
  int v = p[0];
  int d[25];
  for (int k = 0; k < 16; k++) d[k] = v - (int)p[ring[k]];
  for (int k = 16; k < 25; k++) d[k] = d[k - 16];
  int best = -1000;
  for (int s = 0; s < 16; s++) {
    int mn = d[s], mx = d[s];
    for (int k = 1; k < 9; k++) { mn = std::min(mn, d[s + k]); mx = std::max(mx, d[s + k]); }
    best = std::max(best, std::max(mn, -mx));
  }
  return best;
}

// Necessary condition for bam > th (any 9-arc contains one pixel of each opposite pair): lets the CPU baseline
// skip the full arc search on flat pixels, like OpenCV's own early-out.  Never changes a result.
inline bool fast_maybe(const uint8_t* p, const ptrdiff_t* ring, int th) {
  int v = p[0];
  int lo = v - th, hi = v + th;
  int a = p[ring[0]], b = p[ring[8]];
  bool dark = (a < lo) | (b < lo), bright = (a > hi) | (b > hi);
  if (!(dark | bright)) return false;
  a = p[ring[4]]; b = p[ring[12]];
  dark &= (a < lo) | (b < lo); bright &= (a > hi) | (b > hi);
  return dark | bright;
}

void make_ring(size_t pitch, ptrdiff_t ring[16]) {
  for (int k = 0; k < 16; k++) ring[k] = (ptrdiff_t)RING_DY[k] * (ptrdiff_t)pitch + RING_DX[k];
}

struct FastKp { int x, y, score; };

// cv::FAST(img, kps, th, nonmaxSuppression=true) on a w x h view: tested pixels 3 <= x < w-3, 3 <= y < h-3;
// NMS: strictly greater than the 8 neighbours' scores, non-corners and untested pixels count 0; row-major.
void fast_detect(const uint8_t* img, int w, int h, size_t pitch, int th, bool nms, std::vector<FastKp>& out) {
  out.clear();
  if (w < 7 || h < 7) return;
  ptrdiff_t ring[16];
  make_ring(pitch, ring);
  static thread_local std::vector<int16_t> sc;  // -1 = not a corner, else score (bam-1 >= th >= 0)
  sc.assign((size_t)w * h, -1);
  bool any = false;
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      const uint8_t* p = img + (size_t)y * pitch + x;
      if (!fast_maybe(p, ring, th)) continue;
      int bam = fast_bam(p, ring);
      if (bam > th) { sc[(size_t)y * w + x] = (int16_t)(bam - 1); any = true; }
    }
  if (!any) return;
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      int s = sc[(size_t)y * w + x];
      if (s < 0) continue;
      if (nms) {
        bool ok = true;
        for (int dy = -1; dy <= 1 && ok; dy++)
          for (int dx = -1; dx <= 1; dx++) {
            if (!dx && !dy) continue;
            int n = sc[(size_t)(y + dy) * w + (x + dx)];
            if (n < 0) n = 0;  // OpenCV's score rows are zero where there is no corner
            if (n >= s) { ok = false; break; }
          }
        if (!ok) continue;
      }
      out.push_back({x, y, s});
    }
}

// ---------------------------------------------------------------- cv::GaussianBlur 7x7 sigma 2 (App. A.4)
const int GK[7] = {18, 34, 48, 56, 48, 34, 18};
inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}
void gaussian_blur7(const uint8_t* src, int w, int h, size_t spitch, uint8_t* dst, size_t dpitch) {
  // separable: horizontal sums (<= 255*256, fit uint16) for all rows, then the vertical pass with rounding.
  std::vector<uint16_t> tmp((size_t)w * h);
  for (int y = 0; y < h; y++) {
    const uint8_t* s = src + (size_t)y * spitch;
    uint16_t* t = &tmp[(size_t)y * w];
    for (int x = 0; x < w; x++) {
      if (x >= 3 && x < w - 3) {
        t[x] = (uint16_t)(GK[0] * (s[x - 3] + s[x + 3]) + GK[1] * (s[x - 2] + s[x + 2]) + GK[2] * (s[x - 1] + s[x + 1]) +
                          GK[3] * s[x]);
      } else {
        int acc = 0;
        for (int k = -3; k <= 3; k++) acc += GK[k + 3] * s[reflect101(x + k, w)];
        t[x] = (uint16_t)acc;
      }
    }
  }
  for (int y = 0; y < h; y++) {
    const uint16_t* r[7];
    for (int k = -3; k <= 3; k++) r[k + 3] = &tmp[(size_t)reflect101(y + k, h) * w];
    uint8_t* d = dst + (size_t)y * dpitch;
    for (int x = 0; x < w; x++) {
      int acc = GK[0] * (r[0][x] + r[6][x]) + GK[1] * (r[1][x] + r[5][x]) + GK[2] * (r[2][x] + r[4][x]) + GK[3] * r[3][x];
      d[x] = (uint8_t)((acc + 32768) >> 16);
    }
  }
}

// ---------------------------------------------------------------- cv::fastAtan2 (App. A.5); fp32, no FMA
float fast_atan2(float y, float x) {
  const float scale = (float)(180.0 / M_PI);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
                       p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  float ax = std::fabs(x), ay = std::fabs(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)2.2204460492503131e-16);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)2.2204460492503131e-16);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

void make_umax(int umax[HALF_PATCH_SIZE + 2]) {  // ORBextractor.cc:450-469
  int v, v0, vmax = (int)std::floor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
  int vmin = (int)std::ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
  const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
  for (v = 0; v <= HALF_PATCH_SIZE + 1; v++) umax[v] = 0;
  for (v = 0; v <= vmax; ++v) umax[v] = cv_round_d(std::sqrt(hp2 - v * v));
  for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
    while (umax[v0] == umax[v0 + 1]) ++v0;
    umax[v] = v0;
    ++v0;
  }
}

float ic_angle(const uint8_t* img, size_t pitch, int cx, int cy, const int* umax) {  // :77-104
  int m_01 = 0, m_10 = 0;
  const uint8_t* center = img + (size_t)cy * pitch + cx;
  for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
  ptrdiff_t step = (ptrdiff_t)pitch;
  for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
    int v_sum = 0;
    int d = umax[v];
    for (int u = -d; u <= d; ++u) {
      int val_plus = center[u + v * step], val_minus = center[u - v * step];
      v_sum += (val_plus - val_minus);
      m_10 += u * (val_plus + val_minus);
    }
    m_01 += v * v_sum;
  }
  return fast_atan2((float)m_01, (float)m_10);
}

const int8_t PATTERN[1024] = PGB200_ORB_PATTERN_INIT;

// computeOrbDescriptor (:108-147).  cos/sin rule: SURVEY.md App. A.6 -- (float)cos((double)theta).
void orb_descriptor(const uint8_t* img, size_t pitch, int cx, int cy, float angle_deg, uint8_t* desc) {
  const float factorPI = (float)(M_PI / 180.f);
  float angle = angle_deg * factorPI;
  float a = (float)std::cos((double)angle), b = (float)std::sin((double)angle);
  const uint8_t* center = img + (size_t)cy * pitch + cx;
  ptrdiff_t step = (ptrdiff_t)pitch;
  auto get = [&](int idx) -> int {
    float px = (float)PATTERN[2 * idx], py = (float)PATTERN[2 * idx + 1];
    float t0 = px * b, t1 = py * a, t2 = px * a, t3 = py * b;  // built with -ffp-contract=off: no FMA
    int row = cv_round_f(t0 + t1);
    int col = cv_round_f(t2 - t3);
    return center[row * step + col];
  };
  for (int i = 0; i < 32; i++) {
    int val = 0;
    for (int k = 0; k < 8; k++) {
      int t0 = get(16 * i + 2 * k), t1 = get(16 * i + 2 * k + 1);
      val |= (t0 < t1) << k;
    }
    desc[i] = (uint8_t)val;
  }
}

// ---------------------------------------------------------------- octree (App. A.3)
struct Cand { float x, y; int score; };

struct Node {
  int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
  std::vector<int> keys;  // indices into the candidate array, in candidate order
  bool noMore = false;
  long seq = 0;
  std::list<Node>::iterator lit;
};

void divide_node(const Node& n, const std::vector<Cand>& c, Node& n1, Node& n2, Node& n3, Node& n4) {
  const int halfX = (int)std::ceil((float)(n.URx - n.ULx) / 2);
  const int halfY = (int)std::ceil((float)(n.BRy - n.ULy) / 2);
  n1.ULx = n.ULx; n1.ULy = n.ULy; n1.URx = n.ULx + halfX; n1.URy = n.ULy;
  n1.BLx = n.ULx; n1.BLy = n.ULy + halfY; n1.BRx = n.ULx + halfX; n1.BRy = n.ULy + halfY;
  n2.ULx = n1.URx; n2.ULy = n1.URy; n2.URx = n.URx; n2.URy = n.URy;
  n2.BLx = n1.BRx; n2.BLy = n1.BRy; n2.BRx = n.URx; n2.BRy = n.ULy + halfY;
  n3.ULx = n1.BLx; n3.ULy = n1.BLy; n3.URx = n1.BRx; n3.URy = n1.BRy;
  n3.BLx = n.BLx; n3.BLy = n.BLy; n3.BRx = n1.BRx; n3.BRy = n.BLy;
  n4.ULx = n3.URx; n4.ULy = n3.URy; n4.URx = n2.BRx; n4.URy = n2.BRy;
  n4.BLx = n3.BRx; n4.BLy = n3.BRy; n4.BRx = n.BRx; n4.BRy = n.BRy;
  for (int k : n.keys) {
    const Cand& kp = c[k];
    if (kp.x < n1.URx) {
      if (kp.y < n1.BRy) n1.keys.push_back(k); else n3.keys.push_back(k);
    } else if (kp.y < n1.BRy) n2.keys.push_back(k);
    else n4.keys.push_back(k);
  }
  if (n1.keys.size() == 1) n1.noMore = true;
  if (n2.keys.size() == 1) n2.noMore = true;
  if (n3.keys.size() == 1) n3.noMore = true;
  if (n4.keys.size() == 1) n4.noMore = true;
}

typedef std::pair<int, Node*> SizeNode;
bool size_seq_less(const SizeNode& a, const SizeNode& b) {
  if (a.first != b.first) return a.first < b.first;
  return a.second->seq < b.second->seq;  // oracle tie-break rule (reference: pointer value)
}

std::vector<int> distribute_octree(const std::vector<Cand>& c, int minX, int maxX, int minY, int maxY, int N) {
  std::vector<int> result;
  const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
  if (nIni <= 0) return result;  // the reference would divide by zero here
  const float hX = (float)(maxX - minX) / nIni;
  std::list<Node> lNodes;
  std::vector<Node*> vpIniNodes(nIni);
  long seq = 0;
  for (int i = 0; i < nIni; i++) {
    Node ni;
    ni.ULx = (int)(hX * (float)i); ni.ULy = 0;
    ni.URx = (int)(hX * (float)(i + 1)); ni.URy = 0;
    ni.BLx = ni.ULx; ni.BLy = maxY - minY;
    ni.BRx = ni.URx; ni.BRy = maxY - minY;
    ni.seq = seq++;
    lNodes.push_back(ni);
    vpIniNodes[i] = &lNodes.back();
  }
  for (size_t i = 0; i < c.size(); i++) {
    size_t idx = (size_t)(c[i].x / hX);
    if (idx >= (size_t)nIni) idx = nIni - 1;  // cannot happen for x < maxX-minX; guard only
    vpIniNodes[idx]->keys.push_back((int)i);
  }
  auto lit = lNodes.begin();
  while (lit != lNodes.end()) {
    if (lit->keys.size() == 1) { lit->noMore = true; lit++; }
    else if (lit->keys.empty()) lit = lNodes.erase(lit);
    else lit++;
  }
  bool bFinish = false;
  std::vector<SizeNode> vSizeAndPointerToNode;
  auto add_child = [&](Node& n, int* nToExpand) {
    if (n.keys.size() > 0) {
      n.seq = seq++;
      lNodes.push_front(n);
      if (n.keys.size() > 1) {
        if (nToExpand) (*nToExpand)++;
        vSizeAndPointerToNode.push_back(std::make_pair((int)n.keys.size(), &lNodes.front()));
        lNodes.front().lit = lNodes.begin();
      }
    }
  };
  while (!bFinish) {
    int prevSize = (int)lNodes.size();
    lit = lNodes.begin();
    int nToExpand = 0;
    vSizeAndPointerToNode.clear();
    while (lit != lNodes.end()) {
      if (lit->noMore) { lit++; continue; }
      Node n1, n2, n3, n4;
      divide_node(*lit, c, n1, n2, n3, n4);
      add_child(n1, &nToExpand); add_child(n2, &nToExpand); add_child(n3, &nToExpand); add_child(n4, &nToExpand);
      lit = lNodes.erase(lit);
    }
    if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) {
      bFinish = true;
    } else if (((int)lNodes.size() + nToExpand * 3) > N) {
      while (!bFinish) {
        prevSize = (int)lNodes.size();
        std::vector<SizeNode> vPrev = vSizeAndPointerToNode;
        vSizeAndPointerToNode.clear();
        std::sort(vPrev.begin(), vPrev.end(), size_seq_less);
        for (int j = (int)vPrev.size() - 1; j >= 0; j--) {
          Node n1, n2, n3, n4;
          divide_node(*vPrev[j].second, c, n1, n2, n3, n4);
          add_child(n1, nullptr); add_child(n2, nullptr); add_child(n3, nullptr); add_child(n4, nullptr);
          lNodes.erase(vPrev[j].second->lit);
          if ((int)lNodes.size() >= N) break;
        }
        if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) bFinish = true;
      }
    }
  }
  for (auto& n : lNodes) {
    int best = n.keys[0];
    int maxResponse = c[best].score;
    for (size_t k = 1; k < n.keys.size(); k++)
      if (c[n.keys[k]].score > maxResponse) { best = n.keys[k]; maxResponse = c[best].score; }
    result.push_back(best);
  }
  return result;
}

}  // namespace

// ==================================================================== extractor object
struct pgo_orb {
  int nfeatures, nlevels, iniTh, minTh;
  float scaleFactor;
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> nPerLevel;
  int umax[HALF_PATCH_SIZE + 2];
  // products of the last extract call
  std::vector<Image> pyr;
  std::vector<std::vector<Cand>> cands;
  std::vector<std::vector<pgb_keypoint>> lvlKps;  // level coordinates (before rescale)
  double t_stage[6] = {0, 0, 0, 0, 0, 0};
};

extern "C" {

pgo_orb* pgo_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
  if (nfeatures <= 0 || nlevels <= 0 || !(scaleFactor > 1.0f)) return nullptr;
  pgo_orb* o = new pgo_orb;
  o->nfeatures = nfeatures; o->nlevels = nlevels; o->iniTh = iniTh; o->minTh = minTh; o->scaleFactor = scaleFactor;
  // The fork sizes these nlevels+1 (:415-431); entries [0,nlevels) are what anything reads.
  o->scale.resize(nlevels + 1); o->sigma2.resize(nlevels + 1);
  o->invScale.resize(nlevels + 1); o->invSigma2.resize(nlevels + 1);
  o->scale[0] = 1.0f; o->sigma2[0] = 1.0f;
  for (int i = 1; i <= nlevels; i++) {
    o->scale[i] = o->scale[i - 1] * scaleFactor;
    o->sigma2[i] = o->scale[i] * o->scale[i];
  }
  for (int i = 0; i <= nlevels; i++) {
    o->invScale[i] = 1.0f / o->scale[i];
    o->invSigma2[i] = 1.0f / o->sigma2[i];
  }
  o->nPerLevel.resize(nlevels + 1);
  float factor = 1.0f / scaleFactor;
  float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
  int sum = 0;
  for (int level = 0; level < nlevels; level++) {
    o->nPerLevel[level] = cv_round_f(nDesired);
    sum += o->nPerLevel[level];
    nDesired *= factor;
  }
  o->nPerLevel[nlevels] = std::max(nfeatures - sum, 0);
  make_umax(o->umax);
  return o;
}
void pgo_orb_destroy(pgo_orb* o) { delete o; }

int pgo_orb_tables(const pgo_orb* o, float* scale, float* invScale, float* sigma2, float* invSigma2, int32_t* nPer,
                   int32_t* umax16) {
  for (int i = 0; i < o->nlevels; i++) {
    if (scale) scale[i] = o->scale[i];
    if (invScale) invScale[i] = o->invScale[i];
    if (sigma2) sigma2[i] = o->sigma2[i];
    if (invSigma2) invSigma2[i] = o->invSigma2[i];
    if (nPer) nPer[i] = o->nPerLevel[i];
  }
  if (umax16) for (int i = 0; i < 16; i++) umax16[i] = o->umax[i];
  return 0;
}

int pgo_orb_level_size(const pgo_orb* o, int w, int h, int level, int* lw, int* lh) {
  float s = o->invScale[level];
  *lw = cv_round_f((float)w * s);
  *lh = cv_round_f((float)h * s);
  return 0;
}

static double now_s() {
  timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// ORBextractor::operator() (:1042-1104).  Returns the number of keypoints, or -1 if cap is too small.
int pgo_orb_extract(pgo_orb* o, const uint8_t* gray, int w, int h, size_t pitch, pgb_keypoint* kps, uint8_t* desc,
                    int cap) {
  if (w <= 0 || h <= 0) return 0;
  const int L = o->nlevels;
  double t0 = now_s();
  // ComputePyramid (:1106-1131). Borders are never read on this path and are not materialised.
  o->pyr.assign(L, Image());
  for (int level = 0; level < L; level++) {
    int lw, lh;
    pgo_orb_level_size(o, w, h, level, &lw, &lh);
    o->pyr[level] = Image(lw, lh);
    if (level == 0) {
      for (int y = 0; y < h; y++) memcpy(&o->pyr[0].d[(size_t)y * lw], gray + (size_t)y * pitch, lw);
    } else {
      const Image& p = o->pyr[level - 1];
      resize_linear(p.d.data(), p.w, p.h, p.w, o->pyr[level].d.data(), lw, lh, lw);
    }
  }
  double t1 = now_s();
  o->t_stage[0] += t1 - t0;
  // ComputeKeyPointsOctTree (:765-852)
  o->cands.assign(L, {});
  o->lvlKps.assign(L, {});
  const float W = 30;
  std::vector<FastKp> cell;
  for (int level = 0; level < L; level++) {
    double ta = now_s();
    const Image& im = o->pyr[level];
    const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
    const int maxBorderX = im.w - EDGE_THRESHOLD + 3, maxBorderY = im.h - EDGE_THRESHOLD + 3;
    std::vector<Cand>& vToDistribute = o->cands[level];
    const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols <= 0 || nRows <= 0) continue;  // the reference divides by zero for such tiny levels
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    for (int i = 0; i < nRows; i++) {
      const float iniY = (float)(minBorderY + i * hCell);
      float maxY = iniY + hCell + 6;
      if (iniY >= maxBorderY - 3) continue;
      if (maxY > maxBorderY) maxY = (float)maxBorderY;
      for (int j = 0; j < nCols; j++) {
        const float iniX = (float)(minBorderX + j * wCell);
        float maxX = iniX + wCell + 6;
        if (iniX >= maxBorderX - 6) continue;
        if (maxX > maxBorderX) maxX = (float)maxBorderX;
        const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
        const uint8_t* view = im.d.data() + (size_t)y0 * im.w + x0;
        fast_detect(view, cw, ch, im.w, o->iniTh, true, cell);
        if (cell.empty()) fast_detect(view, cw, ch, im.w, o->minTh, true, cell);
        for (const FastKp& k : cell)
          vToDistribute.push_back({(float)(k.x + j * wCell), (float)(k.y + i * hCell), k.score});
      }
    }
    double tb = now_s();
    o->t_stage[1] += tb - ta;
    std::vector<int> keep =
        distribute_octree(vToDistribute, minBorderX, maxBorderX, minBorderY, maxBorderY, o->nPerLevel[level]);
    const int scaledPatchSize = (int)(PATCH_SIZE * o->scale[level]);
    for (int k : keep) {
      pgb_keypoint kp;
      kp.x = vToDistribute[k].x + minBorderX;
      kp.y = vToDistribute[k].y + minBorderY;
      kp.size = (float)scaledPatchSize;
      kp.angle = -1;
      kp.response = (float)vToDistribute[k].score;
      kp.octave = level;
      kp.class_id = -1;
      o->lvlKps[level].push_back(kp);
    }
    o->t_stage[2] += now_s() - tb;
  }
  double t2 = now_s();
  for (int level = 0; level < L; level++)
    for (pgb_keypoint& kp : o->lvlKps[level])
      kp.angle = ic_angle(o->pyr[level].d.data(), o->pyr[level].w, cv_round_f(kp.x), cv_round_f(kp.y), o->umax);
  double t3 = now_s();
  o->t_stage[3] += t3 - t2;
  int total = 0;
  for (int level = 0; level < L; level++) total += (int)o->lvlKps[level].size();
  if (total > cap) return -1;
  int offset = 0;
  Image blurred;
  for (int level = 0; level < L; level++) {
    std::vector<pgb_keypoint>& v = o->lvlKps[level];
    if (v.empty()) continue;
    const Image& im = o->pyr[level];
    double ta = now_s();
    blurred = Image(im.w, im.h);
    gaussian_blur7(im.d.data(), im.w, im.h, im.w, blurred.d.data(), im.w);
    double tb = now_s();
    o->t_stage[4] += tb - ta;
    for (size_t i = 0; i < v.size(); i++) {
      orb_descriptor(blurred.d.data(), im.w, cv_round_f(v[i].x), cv_round_f(v[i].y), v[i].angle,
                     desc + (size_t)(offset + i) * 32);
      pgb_keypoint kp = v[i];
      if (level != 0) { float s = o->scale[level]; kp.x *= s; kp.y *= s; }
      kps[offset + i] = kp;
    }
    o->t_stage[5] += now_s() - tb;
    offset += (int)v.size();
  }
  return total;
}

// stage timers: pyramid, FAST cells, octree, orientation, blur, descriptors (seconds, accumulated)
void pgo_orb_stage_times(pgo_orb* o, double* t6, int reset) {
  for (int i = 0; i < 6; i++) { t6[i] = o->t_stage[i]; if (reset) o->t_stage[i] = 0; }
}

int pgo_orb_get_level(const pgo_orb* o, int level, uint8_t* out, int* w, int* h) {
  if (level < 0 || level >= (int)o->pyr.size()) return -1;
  const Image& im = o->pyr[level];
  *w = im.w; *h = im.h;
  if (out) memcpy(out, im.d.data(), im.d.size());
  return 0;
}
int pgo_orb_get_candidates(const pgo_orb* o, int level, int32_t* xyr, int cap) {
  if (level < 0 || level >= (int)o->cands.size()) return -1;
  const auto& c = o->cands[level];
  if (xyr) {
    if ((int)c.size() > cap) return -1;
    for (size_t i = 0; i < c.size(); i++) { xyr[3 * i] = (int)c[i].x; xyr[3 * i + 1] = (int)c[i].y; xyr[3 * i + 2] = c[i].score; }
  }
  return (int)c.size();
}
int pgo_orb_get_level_keypoints(const pgo_orb* o, int level, pgb_keypoint* out, int cap) {
  if (level < 0 || level >= (int)o->lvlKps.size()) return -1;
  const auto& v = o->lvlKps[level];
  if (out) {
    if ((int)v.size() > cap) return -1;
    memcpy(out, v.data(), v.size() * sizeof(pgb_keypoint));
  }
  return (int)v.size();
}

// ---- primitives, exposed for the cv2 pins and for per-kernel parity tests
void pgo_resize_linear(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
  resize_linear(src, sw, sh, sw, dst, dw, dh, dw);
}
// cv::FAST on a tight w x h image; out xys[cap][3] = x, y, score.  Returns count (or -1 if > cap).
int pgo_fast(const uint8_t* img, int w, int h, int th, int nms, int32_t* xys, int cap) {
  std::vector<FastKp> v;
  fast_detect(img, w, h, w, th, nms != 0, v);
  if ((int)v.size() > cap) return -1;
  for (size_t i = 0; i < v.size(); i++) { xys[3 * i] = v[i].x; xys[3 * i + 1] = v[i].y; xys[3 * i + 2] = v[i].score; }
  return (int)v.size();
}
// Level score map as the CUDA FAST kernel defines it: bam-1 where bam > min_th and the pixel lies in the tested
// region [19, w-19) x [19, h-19) (16-px border + 3-px ring), else 0.
void pgo_fast_score_map(const uint8_t* img, int w, int h, int min_th, uint8_t* out) {
  memset(out, 0, (size_t)w * h);
  ptrdiff_t ring[16];
  make_ring((size_t)w, ring);
  for (int y = EDGE_THRESHOLD; y < h - EDGE_THRESHOLD; y++)
    for (int x = EDGE_THRESHOLD; x < w - EDGE_THRESHOLD; x++) {
      const uint8_t* p = img + (size_t)y * w + x;
      if (!fast_maybe(p, ring, min_th)) continue;
      int bam = fast_bam(p, ring);
      if (bam > min_th) out[(size_t)y * w + x] = (uint8_t)(bam - 1);
    }
}
void pgo_gaussian_blur7(const uint8_t* src, int w, int h, uint8_t* dst) { gaussian_blur7(src, w, h, w, dst, w); }
float pgo_fast_atan2(float y, float x) { return fast_atan2(y, x); }
void pgo_fast_atan2_many(const float* y, const float* x, float* out, int n) {
  for (int i = 0; i < n; i++) out[i] = fast_atan2(y[i], x[i]);
}
float pgo_ic_angle(const uint8_t* img, int w, int h, int cx, int cy) {
  (void)h;
  int umax[HALF_PATCH_SIZE + 2];
  make_umax(umax);
  return ic_angle(img, w, cx, cy, umax);
}
void pgo_orb_descriptor(const uint8_t* blurred, int w, int h, int cx, int cy, float angle_deg, uint8_t* desc32) {
  (void)h;
  orb_descriptor(blurred, w, cx, cy, angle_deg, desc32);
}
// DistributeOctTree on an explicit candidate list (x,y relative to minX/minY); returns kept candidate indices.
int pgo_distribute_octree(const int32_t* xyr, int n, int minX, int maxX, int minY, int maxY, int N, int32_t* keep,
                          int cap) {
  std::vector<Cand> c(n);
  for (int i = 0; i < n; i++) c[i] = {(float)xyr[3 * i], (float)xyr[3 * i + 1], xyr[3 * i + 2]};
  std::vector<int> k = distribute_octree(c, minX, maxX, minY, maxY, N);
  if ((int)k.size() > cap) return -1;
  for (size_t i = 0; i < k.size(); i++) keep[i] = k[i];
  return (int)k.size();
}

}  // extern "C"

// cv::flip + cvtColor(8U, RGB/BGR[A] -> GRAY) (image_sequence_reader.cc:163-175, Tracking.cc:243-258).
// formula 0: OpenCV 2.4 (yuv_shift 14: R2Y 4899, G2Y 9617, B2Y 1868), formula 1: OpenCV >= 3 (shift 15: 9798, 19235, 3735).
extern "C" void pgo_to_gray(const uint8_t* src, int w, int h, int channels, int rgb_order, int vflip, int hflip, int formula,
                            uint8_t* dst) {
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const uint8_t* p = src + ((size_t)(vflip ? h - 1 - y : y) * w + (hflip ? w - 1 - x : x)) * channels;
      int v = p[0];
      if (channels > 1) {
        const int r = rgb_order ? p[0] : p[2], g = p[1], b = rgb_order ? p[2] : p[0];
        v = formula ? (r * 9798 + g * 19235 + b * 3735 + 16384) >> 15 : (r * 4899 + g * 9617 + b * 1868 + 8192) >> 14;
      }
      dst[(size_t)y * w + x] = (uint8_t)v;
    }
}
