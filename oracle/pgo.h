/* ORACLE -- TEST INFRASTRUCTURE ONLY (see pgo_orb.cc header).  C API of liboracle.so, loaded by tests via ctypes. */
#ifndef PGO_H_
#define PGO_H_
#include <stddef.h>
#include <stdint.h>
#include "../include/pgb200.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct pgo_orb pgo_orb;
pgo_orb* pgo_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh);
void pgo_orb_destroy(pgo_orb*);
int pgo_orb_tables(const pgo_orb*, float* scale, float* invScale, float* sigma2, float* invSigma2, int32_t* nPer,
                   int32_t* umax16);
int pgo_orb_level_size(const pgo_orb*, int w, int h, int level, int* lw, int* lh);
int pgo_orb_extract(pgo_orb*, const uint8_t* gray, int w, int h, size_t pitch, pgb_keypoint* kps, uint8_t* desc,
                    int cap);
void pgo_orb_stage_times(pgo_orb*, double* t6, int reset);
int pgo_orb_get_level(const pgo_orb*, int level, uint8_t* out, int* w, int* h);
int pgo_orb_get_candidates(const pgo_orb*, int level, int32_t* xyr, int cap);
int pgo_orb_get_level_keypoints(const pgo_orb*, int level, pgb_keypoint* out, int cap);
void pgo_resize_linear(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
int pgo_fast(const uint8_t* img, int w, int h, int th, int nms, int32_t* xys, int cap);
void pgo_fast_score_map(const uint8_t* img, int w, int h, int min_th, uint8_t* out);
void pgo_gaussian_blur7(const uint8_t* src, int w, int h, uint8_t* dst);
float pgo_fast_atan2(float y, float x);
void pgo_fast_atan2_many(const float* y, const float* x, float* out, int n);
float pgo_ic_angle(const uint8_t* img, int w, int h, int cx, int cy);
void pgo_orb_descriptor(const uint8_t* blurred, int w, int h, int cx, int cy, float angle_deg, uint8_t* desc32);
int pgo_distribute_octree(const int32_t* xyr, int n, int minX, int maxX, int minY, int maxY, int N, int32_t* keep,
                          int cap);
int pgo_descriptor_distance(const uint8_t* a, const uint8_t* b);
int pgo_search_by_projection(const pgb_keypoint* cur_kps, const uint8_t* cur_desc, int n_cur, const float* q_uv,
                             const int32_t* q_octave, const float* q_angle, const uint8_t* q_desc,
                             const uint8_t* q_valid, int n_q, float minX, float maxX, float minY, float maxY, float th,
                             const float* scale_factors, int nlevels, int check_ori, int32_t* match_of_cur,
                             int32_t* best_dist_of_q);
int pgo_match_consecutive(const pgb_keypoint* prev_kps, const uint8_t* prev_desc, int n_prev,
                          const pgb_keypoint* cur_kps, const uint8_t* cur_desc, int n_cur, float flow_x, float flow_y,
                          float maxX, float maxY, float th, const float* scale_factors, int nlevels,
                          int32_t* match_of_cur);
double pgo_bench_extract_match(const uint8_t* frames, int n, int w, int h, const float* flow_xy, int nfeatures,
                               float scale, int nlevels, int iniTh, int minTh, float th, int nthreads,
                               int64_t* total_kps, int64_t* total_matches);
#ifdef __cplusplus
}
#endif
#endif
