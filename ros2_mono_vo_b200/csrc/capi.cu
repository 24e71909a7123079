// capi.cu -- extern "C" entry points of libmonovo_b200.so (context, ORB, kNN).  See include/monovo_b200.h.
#include "context.cuh"
#include "host_hash.hpp"
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <mutex>

static std::string g_create_error;
static std::mutex g_create_mutex;

using namespace mvo;

#define MVO_CHECK_ARG(ctx, cond, msg)  \
  do {                                 \
    if (!(cond)) {                     \
      (ctx)->set_error(msg);           \
      return MVO_ERR_INVALID;          \
    }                                  \
  } while (0)

static int ensure_stage(mvo_ctx* c, size_t bytes) {
  if (bytes > c->h_stage.n) MVO_CUDA_TRY(c, c->h_stage.alloc(bytes + bytes / 4 + 4096));
  return MVO_OK;
}

namespace mvo {
// points3d_to_pointcloud_msg's loop (src/utils.cpp:225-241): (x, y, z) camera / OpenCV axes -> ROS axes (z, -x, -y),
// three floats per point; one thread per float of the output (coalesced 4-byte accesses on both sides)
__global__ void __launch_bounds__(256) pack_cloud_kernel(const float* __restrict__ xyz, int n, float* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3 * n) return;
  const int p = e / 3, k = e - 3 * p;
  // out[0] = z, out[1] = -x, out[2] = -y
  const float v = xyz[3 * p + (k == 0 ? 2 : k - 1)];
  out[e] = k == 0 ? v : -v;
}
}  // namespace mvo

extern "C" {

const char* mvo_version(void) { return "monovo_b200 0.1 (sm_100a)"; }

int mvo_create(mvo_ctx** out, const mvo_config* cfg) {
  std::lock_guard<std::mutex> lock(g_create_mutex);
  if (!out || !cfg) {
    g_create_error = "mvo_create: null argument";
    return MVO_ERR_INVALID;
  }
  *out = nullptr;
  if (cfg->max_width < 16 || cfg->max_height < 16 || cfg->nfeatures < 1 || cfg->batch < 1 ||
      cfg->max_width > 16384 || cfg->max_height > 16384) {
    g_create_error = "mvo_create: configuration out of range";
    return MVO_ERR_INVALID;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    g_create_error = std::string("mvo_create: no CUDA device (") + cudaGetErrorString(e) +
                     "); libmonovo_b200 has no CPU fallback";
    return MVO_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) {
    g_create_error = "mvo_create: device ordinal out of range";
    return MVO_ERR_INVALID;
  }
  e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) {
    g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return MVO_ERR_CUDA;
  }
  mvo_ctx* c = new mvo_ctx();
  c->cfg = *cfg;
  if (c->cfg.max_points <= 0) c->cfg.max_points = 2 * c->cfg.nfeatures;
  if (c->cfg.ransac_seed == 0) c->cfg.ransac_seed = 0xFFFFFFFFFFFFFFFFull;
  if (const char* e = getenv("MVO_LK_IMPL")) c->dbg_lk_impl = atoi(e);      // A/B aid: see mvo_debug_set
  if (cfg->cuda_stream) {
    c->stream = (cudaStream_t)cfg->cuda_stream;
  } else {
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
      delete c;
      return MVO_ERR_CUDA;
    }
    c->own_stream = true;
  }
  c->main_stream = c->stream;
  {
    // the model searches (F, E, H + tail) are chains of small latency-bound kernels: with the highest priority their
    // CTAs are placed first while the next step's ORB / LK kernels (main stream) fill the rest of the GPU
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    cudaError_t first = cudaSuccess;
    auto keep = [&first](cudaError_t r) {
      if (first == cudaSuccess && r != cudaSuccess) first = r;
    };
    for (int k = 0; k < 4; ++k)
      keep(cudaStreamCreateWithPriority(&c->aux_stream[k], cudaStreamNonBlocking, k == 2 ? prio_lo : prio_hi));
    keep(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    keep(cudaStreamCreateWithPriority(&c->lk_stream, cudaStreamNonBlocking, prio_lo));
    keep(cudaEventCreateWithFlags(&c->ev_unpack, cudaEventDisableTiming));
    keep(cudaEventCreateWithFlags(&c->ev_lk_done, cudaEventDisableTiming));
    keep(cudaEventCreateWithFlags(&c->ev_graph_done, cudaEventDisableTiming));
    keep(cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking));
    for (auto& sl : c->slots) {
      keep(cudaEventCreateWithFlags(&sl.ev_up, cudaEventDisableTiming));
      keep(cudaEventCreateWithFlags(&sl.ev_free, cudaEventDisableTiming));
      keep(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
    }
    for (auto& ev : c->ev_fork) keep(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : c->ev_join) keep(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    keep(cudaEventCreateWithFlags(&c->ev_tail, cudaEventDisableTiming));
    for (cudaEvent_t* ev : {&c->ev_o_orb, &c->ev_o_lk, &c->ev_out_orb, &c->ev_out_lk})
      keep(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    for (auto& t : c->timers) {
      keep(cudaEventCreate(&t.beg));
      keep(cudaEventCreate(&t.end));
    }
    if (first != cudaSuccess) {   // a NULL aux stream would silently be the legacy default stream
      g_create_error = std::string("mvo_create: stream / event creation failed: ") + cudaGetErrorString(first);
      mvo_destroy(c);
      return MVO_ERR_CUDA;
    }
  }
  *out = c;
  return MVO_OK;
}

void mvo_destroy(mvo_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  c->stream = c->main_stream;
  cudaStreamSynchronize(c->stream);
  for (auto& st : c->aux_stream)
    if (st) {
      cudaStreamSynchronize(st);
      cudaStreamDestroy(st);
    }
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
  }
  if (c->out_stream) {
    cudaStreamSynchronize(c->out_stream);
    cudaStreamDestroy(c->out_stream);
  }
  for (cudaEvent_t ev : {c->ev_o_orb, c->ev_o_lk, c->ev_out_orb, c->ev_out_lk})
    if (ev) cudaEventDestroy(ev);
  for (auto& sl : c->slots) {
    if (sl.ev_up) cudaEventDestroy(sl.ev_up);
    if (sl.ev_free) cudaEventDestroy(sl.ev_free);
    if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    sl.stage.release();
    sl.h_res.release();
    sl.h_flags.release();
    sl.h_occ.release();
  }
  for (auto& ev : c->ev_fork)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->ev_join)
    if (ev) cudaEventDestroy(ev);
  if (c->ev_tail) cudaEventDestroy(c->ev_tail);
  if (c->lk_stream) {
    cudaStreamSynchronize(c->lk_stream);
    cudaStreamDestroy(c->lk_stream);
  }
  if (c->ev_unpack) cudaEventDestroy(c->ev_unpack);
  if (c->ev_lk_done) cudaEventDestroy(c->ev_lk_done);
  if (c->ev_graph_done) cudaEventDestroy(c->ev_graph_done);
  for (auto& gph : c->step_graph)
    if (gph.exec) cudaGraphExecDestroy(gph.exec);
  c->img_in.release(); c->pyr.release(); c->blur.release(); c->xtab.release(); c->ytab.release();
  c->cand_xy.release(); c->cand_score.release(); c->cand_count.release(); c->cand_sel.release(); c->sel_count.release(); c->hist.release();
  c->c2_key.release(); c->c2_key_sorted.release(); c->c2_ra.release(); c->c2_ra_sorted.release();
  c->c2_count.release(); c->kps.release(); c->desc.release(); c->kp_valid.release(); c->kp_count.release();
  c->flags.release(); c->occ.release(); c->cloud.release(); c->pack_tmp.release(); c->h_stage.release(); c->prev_kps.release(); c->prev_desc.release();
  c->prev_kp_count.release(); c->kp_xy.release(); c->prev_kp_xy.release(); c->d_results.release(); c->knn_q.release(); c->knn_t.release(); c->knn_best.release();
  c->knn_matches.release(); c->knn_nmatch.release(); c->knn_counts.release();
  c->lk_pyr[0].release(); c->lk_pyr[1].release(); c->lk_pts_in.release(); c->lk_pts_out.release();
  c->lk_status.release(); c->lk_err.release(); c->lk_npts.release(); c->lk_work.release();
  c->rs.release();
  c->pnp.release();
  for (auto& t : c->timers) {
    if (t.beg) cudaEventDestroy(t.beg);
    if (t.end) cudaEventDestroy(t.end);
  }
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* mvo_last_error(const mvo_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }
void* mvo_cuda_stream(mvo_ctx* c) { return c ? (void*)c->main_stream : nullptr; }
int mvo_batch(const mvo_ctx* c) { return c ? c->cfg.batch : 0; }
uint64_t mvo_launch_count(const mvo_ctx* c) { return c ? c->launches : 0; }
int mvo_orb_num_levels(void) { return kLevels; }

static int desc_cache_put_device(mvo_ctx* c, const uint8_t* host_copy, const uint8_t* dev, int n);

// ------------------------------------------------------------------------------------------------
int mvo_orb_detect_and_compute(mvo_ctx* c, const uint8_t* img, int w, int h, int stride, int channels,
                               mvo_keypoint* kps, uint8_t* desc, int cap, int* n_out) {
  if (!c) return MVO_ERR_INVALID;
  MVO_REQUIRE_IDLE(c);
  MVO_CHECK_ARG(c, img && kps && n_out && cap >= 0, "mvo_orb_detect_and_compute: null argument");
  MVO_CHECK_ARG(c, channels == 1 || channels == 3, "channels must be 1 or 3");
  MVO_CHECK_ARG(c, stride >= w * channels, "stride smaller than a row");
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  int rc = orb_prepare(c, w, h);
  if (rc) return rc;
  const OrbGeom& g = c->geom;
  // the single-image API drives stream 0 of the group; replicate the frame for the other streams
  if (g.batch != 1) {
    c->set_error("mvo_orb_detect_and_compute needs a batch==1 context (use mvo_group_step for groups)");
    return MVO_ERR_INVALID;
  }
  size_t kb = 0, db = 0;
  uint8_t* hs = nullptr;
  int n = 0, flags = 0;
  for (;;) {
    rc = orb_upload(c, img, w, h, stride, channels, 0);
    if (rc) return rc;
    rc = orb_run_detect(c, desc != nullptr);
    if (rc) return rc;
    kb = (size_t)g.kp_cap * sizeof(mvo_keypoint);
    db = (size_t)g.kp_cap * 32;
    rc = ensure_stage(c, kb + db + 64);
    if (rc) return rc;
    hs = c->h_stage.p;
    MVO_CUDA_TRY(c, cudaMemcpyAsync(hs, c->kp_count.p, 4, cudaMemcpyDeviceToHost, c->stream));
    MVO_CUDA_TRY(c, cudaMemcpyAsync(hs + 16, c->flags.p, 4, cudaMemcpyDeviceToHost, c->stream));
    MVO_CUDA_TRY(c, cudaMemcpyAsync(hs + 32, c->occ.p, 8, cudaMemcpyDeviceToHost, c->stream));
    MVO_CUDA_TRY(c, cudaMemcpyAsync(hs + 64, c->kps.p, kb, cudaMemcpyDeviceToHost, c->stream));
    if (desc) MVO_CUDA_TRY(c, cudaMemcpyAsync(hs + 64 + kb, c->desc.p, db, cudaMemcpyDeviceToHost, c->stream));
    MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    n = *reinterpret_cast<int*>(hs);
    flags = *reinterpret_cast<int*>(hs + 16);
    if (!(flags & 1) || c->cand_scale >= 16) break;
    // FAST candidate list overflow (frames that are mostly noise: ~10 % of the pixels are corners): cv::ORB still
    // returns features there, so the list capacity is doubled and the frame is run again instead of failing
    c->cand_scale *= 2;
    c->geom_w = c->geom_h = -1;                  // rebuild the geometry with the larger lists
    c->have_prev = false;
    rc = orb_prepare(c, w, h);
    if (rc) return rc;
  }
  c->occ_single[0] = reinterpret_cast<int*>(hs + 32)[0];
  c->occ_single[1] = reinterpret_cast<int*>(hs + 32)[1];
  c->occ_from_group = false;
  if (flags & 1) {
    c->set_error("FAST candidate list overflow");
    return MVO_ERR_CAPACITY;
  }
  if ((flags & 2) || n > cap) {
    *n_out = n;
    c->set_error("keypoint capacity exceeded");
    return MVO_ERR_CAPACITY;
  }
  memcpy(kps, hs + 64, (size_t)n * sizeof(mvo_keypoint));
  if (desc) memcpy(desc, hs + 64 + kb, (size_t)n * 32);
  *n_out = n;
  // the descriptors stay on the device under their content hash: a find_matches on this block does not upload it again
  if (desc) return desc_cache_put_device(c, desc, c->desc.p, n);
  return MVO_OK;
}

int mvo_orb_compute(mvo_ctx* c, const uint8_t* img, int w, int h, int stride, int channels,
                    const mvo_keypoint* kps_in, int n, uint8_t* desc, uint8_t* valid) {
  if (!c) return MVO_ERR_INVALID;
  MVO_REQUIRE_IDLE(c);
  MVO_CHECK_ARG(c, img && kps_in && desc && n >= 0, "mvo_orb_compute: null argument");
  MVO_CHECK_ARG(c, channels == 1 || channels == 3, "channels must be 1 or 3");
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  int rc = orb_prepare(c, w, h);
  if (rc) return rc;
  const OrbGeom& g = c->geom;
  MVO_CHECK_ARG(c, g.batch == 1, "mvo_orb_compute needs a batch==1 context");
  if (n > g.kp_cap) {
    c->set_error("more keypoints than the context capacity");
    return MVO_ERR_CAPACITY;
  }
  rc = orb_upload(c, img, w, h, stride, channels, 0);
  if (rc) return rc;
  rc = orb_run_levels_only(c);
  if (rc) return rc;
  if (n == 0) {
    MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return MVO_OK;
  }
  MVO_CUDA_TRY(c, cudaMemcpyAsync(c->kps.p, kps_in, (size_t)n * sizeof(mvo_keypoint), cudaMemcpyHostToDevice, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(c->kp_count.p, &n, 4, cudaMemcpyHostToDevice, c->stream));
  rc = orb_run_brief_given(c, n);
  if (rc) return rc;
  MVO_CUDA_TRY(c, cudaMemcpyAsync(desc, c->desc.p, (size_t)n * 32, cudaMemcpyDeviceToHost, c->stream));
  if (valid) MVO_CUDA_TRY(c, cudaMemcpyAsync(valid, c->kp_valid.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return MVO_OK;
}

int mvo_orb_level_size(mvo_ctx* c, int level, int* w, int* h) {
  if (!c || level < 0 || level >= kLevels || c->geom_w < 0) return MVO_ERR_INVALID;
  if (w) *w = c->geom.lv[level].w;
  if (h) *h = c->geom.lv[level].h;
  return MVO_OK;
}

int mvo_orb_get_level(mvo_ctx* c, int level, int blurred, uint8_t* out, int out_stride) {
  if (!c || !out || level < 0 || level >= kLevels || c->geom_w < 0) return MVO_ERR_INVALID;
  const LevelGeom& lv = c->geom.lv[level];
  if (out_stride < lv.w) return MVO_ERR_INVALID;
  const uint8_t* src = (blurred ? c->blur.p : c->pyr.p) + lv.off;
  MVO_CUDA_TRY(c, cudaMemcpy2DAsync(out, out_stride, src, lv.pitch, lv.w, lv.h, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return MVO_OK;
}

int mvo_orb_get_fast(mvo_ctx* c, int level, uint32_t* xy, int32_t* score, int cap, int* n_out) {
  if (!c || !xy || !score || !n_out || level < 0 || level >= kLevels || c->geom_w < 0) return MVO_ERR_INVALID;
  const LevelGeom& lv = c->geom.lv[level];
  int n = 0;
  MVO_CUDA_TRY(c, cudaMemcpyAsync(&n, c->cand_count.p + level, 4, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  n = std::min(n, lv.cand_cap);
  *n_out = n;
  if (n > cap) return MVO_ERR_CAPACITY;
  MVO_CUDA_TRY(c, cudaMemcpyAsync(xy, c->cand_xy.p + lv.cand_off, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(score, c->cand_score.p + lv.cand_off, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return MVO_OK;
}

// ------------------------------------------------------------------------------------------------
// ---- descriptor cache: content-keyed device copies of N x 32 blocks (least recently used of kDescCache replaced) ----
static DescCacheEntry* desc_cache_find(mvo_ctx* c, uint64_t hash, int n) {
  for (auto& e : c->dcache)
    if (e.hash == hash && e.n == n && hash) {
      e.stamp = ++c->dcache_clock;
      return &e;
    }
  return nullptr;
}
static DescCacheEntry* desc_cache_victim(mvo_ctx* c) {
  DescCacheEntry* v = &c->dcache[0];
  for (auto& e : c->dcache)
    if (e.stamp < v->stamp) v = &e;
  return v;
}
// the block produced by detect_and_compute stays on the device under the hash of the copy handed to the caller
static int desc_cache_put_device(mvo_ctx* c, const uint8_t* host_copy, const uint8_t* dev, int n) {
  if (!c->cache_enabled || n <= 0) return MVO_OK;
  const uint64_t hash = content_hash(host_copy, (size_t)n * 32);
  if (desc_cache_find(c, hash, n)) return MVO_OK;
  DescCacheEntry* e = desc_cache_victim(c);
  e->hash = 0;
  MVO_CUDA_TRY(c, e->buf.alloc((size_t)n * 32));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(e->buf.p, dev, (size_t)n * 32, cudaMemcpyDeviceToDevice, c->stream));
  e->hash = hash;
  e->n = n;
  e->stamp = ++c->dcache_clock;
  return MVO_OK;
}
// device pointer of a host descriptor block: the cached copy when its content is known, else upload (and remember)
static int desc_cache_get(mvo_ctx* c, const uint8_t* host, int n, const uint8_t** dev, mvo::DevBuf<uint8_t>& fallback) {
  if (n <= 0) {
    MVO_CUDA_TRY(c, fallback.alloc(32));
    *dev = fallback.p;
    return MVO_OK;
  }
  if (!c->cache_enabled) {
    MVO_CUDA_TRY(c, fallback.alloc((size_t)n * 32));
    MVO_CUDA_TRY(c, cudaMemcpyAsync(fallback.p, host, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
    *dev = fallback.p;
    return MVO_OK;
  }
  const uint64_t hash = content_hash(host, (size_t)n * 32);
  if (DescCacheEntry* e = desc_cache_find(c, hash, n)) {
    c->cache_stats[0]++;
    *dev = e->buf.p;
    return MVO_OK;
  }
  c->cache_stats[1]++;
  DescCacheEntry* e = desc_cache_victim(c);
  e->hash = 0;
  MVO_CUDA_TRY(c, e->buf.alloc((size_t)n * 32));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(e->buf.p, host, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
  e->hash = hash;
  e->n = n;
  e->stamp = ++c->dcache_clock;
  *dev = e->buf.p;
  return MVO_OK;
}

static int knn_host(mvo_ctx* c, const uint8_t* q, int nq, const uint8_t* t, int nt, double ratio) {
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  MVO_CUDA_TRY(c, c->knn_counts.alloc(2));
  const uint8_t *qd = nullptr, *td = nullptr;
  int rc = desc_cache_get(c, q, nq, &qd, c->knn_q);
  if (rc) return rc;
  rc = desc_cache_get(c, t, nt, &td, c->knn_t);     // the query entry was just stamped: it is not the victim
  if (rc) return rc;
  rc = ensure_stage(c, 64);
  if (rc) return rc;
  if (rc) return rc;
  int* hc = reinterpret_cast<int*>(c->h_stage.p);
  hc[0] = nq;
  hc[1] = nt;
  MVO_CUDA_TRY(c, cudaMemcpyAsync(c->knn_counts.p, hc, 8, cudaMemcpyHostToDevice, c->stream));
  return knn_run(c, qd, c->knn_counts.p, std::max(nq, 1), std::max(nq, 1), td, c->knn_counts.p + 1,
                 std::max(nt, 1), std::max(nt, 1), ratio, 1);
}

int mvo_knn_ratio(mvo_ctx* c, const uint8_t* q, int nq, const uint8_t* t, int nt, double ratio, mvo_dmatch* out,
                  int* n_out) {
  if (!c) return MVO_ERR_INVALID;
  MVO_REQUIRE_IDLE(c);
  MVO_CHECK_ARG(c, n_out && nq >= 0 && nt >= 0 && (nq == 0 || (q && out)) && (nt == 0 || t),
                "mvo_knn_ratio: bad argument");
  *n_out = 0;
  if (nq == 0) return MVO_OK;
  int rc = knn_host(c, q, nq, t, nt, ratio);
  if (rc) return rc;
  rc = ensure_stage(c, (size_t)nq * sizeof(mvo_dmatch) + 64);
  if (rc) return rc;
  uint8_t* hs = c->h_stage.p;
  MVO_CUDA_TRY(c, cudaMemcpyAsync(hs, c->knn_nmatch.p, 4, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(hs + 64, c->knn_matches.p, (size_t)nq * sizeof(mvo_dmatch), cudaMemcpyDeviceToHost,
                                  c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  const int n = *reinterpret_cast<int*>(hs);
  memcpy(out, hs + 64, (size_t)n * sizeof(mvo_dmatch));
  *n_out = n;
  return MVO_OK;
}

int mvo_knn2(mvo_ctx* c, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist) {
  if (!c) return MVO_ERR_INVALID;
  MVO_REQUIRE_IDLE(c);
  MVO_CHECK_ARG(c, nq >= 0 && nt >= 0 && (nq == 0 || (q && idx && dist)) && (nt == 0 || t), "mvo_knn2: bad argument");
  if (nq == 0) return MVO_OK;
  int rc = knn_host(c, q, nq, t, nt, 0.7);
  if (rc) return rc;
  rc = ensure_stage(c, (size_t)nq * 8);
  if (rc) return rc;
  uint32_t* hb = reinterpret_cast<uint32_t*>(c->h_stage.p);
  MVO_CUDA_TRY(c, cudaMemcpyAsync(hb, c->knn_best.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < nq; ++i)
    for (int k = 0; k < 2; ++k) {
      const uint32_t key = hb[2 * i + k];
      idx[2 * i + k] = key == 0xFFFFFFFFu ? -1 : (int)(key & 0x3FFFFFu);
      dist[2 * i + k] = key == 0xFFFFFFFFu ? -1 : (int)(key >> 22);
    }
  return MVO_OK;
}

// ---- cache statistics ------------------------------------------------------------------
int mvo_cache_stats(mvo_ctx* c, uint64_t stats[4]) {
  if (!c || !stats) return MVO_ERR_INVALID;
  for (int i = 0; i < 4; ++i) stats[i] = c->cache_stats[i];
  return MVO_OK;
}

// ---- SURVEY 8(f) #4 -------------------------------------------------------------------------------
int mvo_set_occupancy_grid(mvo_ctx* c, int grid_div) {
  if (!c) return MVO_ERR_INVALID;
  MVO_REQUIRE_IDLE(c);
  MVO_CHECK_ARG(c, grid_div >= 0, "mvo_set_occupancy_grid: negative cell size");
  c->occupancy_div = grid_div;
  if (c->geom_w > 0) orb_set_grid(c, c->geom_w, c->geom_h);
  return MVO_OK;
}

int mvo_orb_occupancy(mvo_ctx* c, int stream, int* occupied_cells, int* total_cells) {
  if (!c) return MVO_ERR_INVALID;
  MVO_CHECK_ARG(c, occupied_cells && total_cells && stream >= 0 && stream < c->cfg.batch, "mvo_orb_occupancy: bad argument");
  if (c->occ_from_group) {
    MVO_CHECK_ARG(c, c->out_slot >= 0 && c->slots[c->out_slot].h_occ.p, "mvo_orb_occupancy: no finished group step");
    *occupied_cells = c->slots[c->out_slot].h_occ.p[2 * stream];
    *total_cells = c->slots[c->out_slot].h_occ.p[2 * stream + 1];
  } else {
    MVO_CHECK_ARG(c, stream == 0, "mvo_orb_occupancy: single-call contexts have one stream");
    *occupied_cells = c->occ_single[0];
    *total_cells = c->occ_single[1];
  }
  if (*occupied_cells < 0) {
    c->set_error("the keypoint-distribution grid is switched off (mvo_set_occupancy_grid) or does not fit (rows x cols > 8192)");
    return MVO_ERR_UNSUPPORTED;
  }
  return MVO_OK;
}

int mvo_pack_pointcloud(mvo_ctx* c, const float* points_xyz, int n, int points_on_device, uint8_t* out, int out_on_device) {
  if (!c) return MVO_ERR_INVALID;
  MVO_REQUIRE_IDLE(c);
  MVO_CHECK_ARG(c, n >= 0 && (n == 0 || (points_xyz && out)), "mvo_pack_pointcloud: bad argument");
  if (n == 0) return MVO_OK;
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  const size_t bytes = (size_t)n * 12;
  MVO_CUDA_TRY(c, c->pack_tmp.alloc(2 * bytes));
  const float* src = points_xyz;
  if (!points_on_device) {
    MVO_CUDA_TRY(c, cudaMemcpyAsync(c->pack_tmp.p, points_xyz, bytes, cudaMemcpyHostToDevice, c->stream));
    src = reinterpret_cast<const float*>(c->pack_tmp.p);
  }
  float* dst = out_on_device ? reinterpret_cast<float*>(out) : reinterpret_cast<float*>(c->pack_tmp.p + bytes);
  pack_cloud_kernel<<<(3 * n + 255) / 256, 256, 0, c->stream>>>(src, n, dst);
  c->launches++;
  MVO_CUDA_TRY(c, cudaGetLastError());
  if (!out_on_device) MVO_CUDA_TRY(c, cudaMemcpyAsync(out, dst, bytes, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return MVO_OK;
}

}  // extern "C"
