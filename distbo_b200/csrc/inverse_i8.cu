// The inverse products of the fit on the tcgen05 digit-plane kernel (digit_planes.cuh):
//   distbo_lauum_i8_f64 : K^-1 = W^T W from the planes of T = W^T (split per column of W);
//   distbo_trtri_i8_f64 : the top levels of W = L^-1,  W21 = -W22 (L21 W11), two products per level.
// Reference: autodiff through tf.cholesky / tf.matrix_triangular_solve, train/log_gaus_likelihood.py:119-125,
//            183-184, 211-214.
#include "../../include/distbo_b200.h"
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "digit_planes.cuh"

using namespace db;

// ---- LAUUM on the tcgen05 tensor cores: Kinv = W^T W = T T^T with T = W^T in int8 digit planes ---------------
namespace db {
static int lauum_i8_items(int n_pad, std::vector<int4>* items, std::vector<int>* off) {
  const int sms = device_sm_count();
  const int nb = n_pad / I8_TILE, nk = n_pad / I8_KC;
  struct It { int4 w; double cost; };
  std::vector<It> all;
  for (int ib = 0; ib < nb; ++ib)
    for (int jb = 0; jb <= 2 * ib + 1; ++jb)
      all.push_back({make_int4(ib * I8_TILE, jb * I8_RB, 2 * ib, nk), (double)(nk - 2 * ib) + 2.0});
  // longest first onto the least loaded CTA (items are generated with decreasing k range already)
  std::vector<double> load(sms, 0.0);
  std::vector<std::vector<int4>> lists(sms);
  for (const It& it : all) {
    int best = 0;
    for (int c = 1; c < sms; ++c)
      if (load[c] < load[best]) best = c;
    load[best] += it.cost;
    lists[best].push_back(it.w);
  }
  if (items) {
    items->clear();
    off->assign(sms + 1, 0);
    for (int c = 0; c < sms; ++c) {
      (*off)[c] = (int)items->size();
      items->insert(items->end(), lists[c].begin(), lists[c].end());
    }
    (*off)[sms] = (int)items->size();
  }
  return (int)all.size();
}
}  // namespace db

extern "C" int64_t distbo_lauum_i8_sched_words(int n_pad) {
  if (n_pad <= 0 || n_pad % 128) return 0;
  return (int64_t)lauum_i8_items(n_pad, nullptr, nullptr) * 4 + device_sm_count() + 1 + 3;
}

extern "C" int distbo_lauum_i8_plan(int n_pad, int32_t* sched) {
  if (!sched || n_pad <= 0 || n_pad % 128 || n_pad / I8_RB > I8_MAX_RB) return DISTBO_EARG;
  std::vector<int4> items;
  std::vector<int> off;
  lauum_i8_items(n_pad, &items, &off);
  // layout: [sms + 1 offsets, padded to a multiple of 4 words][items]
  const size_t off_words = (off.size() + 3) / 4 * 4;
  DB_CUDA(cudaMemcpy(sched, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice));
  DB_CUDA(cudaMemcpy(sched + off_words, items.data(), items.size() * sizeof(int4), cudaMemcpyHostToDevice));
  return DISTBO_OK;
}

extern "C" int64_t distbo_lauum_i8_workspace(int n_pad) {
  return (int64_t)n_pad * n_pad + n_pad;   // doubles: 8 digit planes of W^T (8 bytes per element) + column exponents
}

extern "C" int distbo_lauum_i8_f64(const double* W, int n_pad, const int32_t* sched, double* ws, double* Kinv,
                                   void* stream) {
  if (!W || !sched || !ws || !Kinv || n_pad <= 0 || n_pad % 128 || n_pad / I8_RB > I8_MAX_RB) return DISTBO_EARG;
  cudaStream_t st = (cudaStream_t)stream;
  static PerDeviceOnce once;
  if (once.needed()) {
    DB_CUDA(cudaFuncSetAttribute(score_i8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM));
    once.mark();
  }
  signed char* T8 = reinterpret_cast<signed char*>(ws);
  double* colscale = ws + (size_t)n_pad * n_pad;
  DB_CUDA(cudaMemsetAsync(colscale, 0, sizeof(double) * n_pad, st));
  i8_colmax_kernel<<<dim3(n_pad / 64, (n_pad + 511) / 512), 256, 0, st>>>(
      W, n_pad, reinterpret_cast<unsigned long long*>(colscale), nullptr, 0);
  DB_CHECK_LAUNCH();
  i8_colscale_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(colscale, n_pad);
  DB_CHECK_LAUNCH();
  const int nb64 = n_pad / 64;
  i8_wtsplit_kernel<<<nb64 * (nb64 + 1) / 2, 256, 0, st>>>(W, n_pad, colscale, T8, nullptr, 0);
  DB_CHECK_LAUNCH();
  CUtensorMap mapA, mapAhi, mapB;
  if (!tma::encode_u8_planes(&mapA, T8, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_TILE, I8_A_LO) ||
      !tma::encode_u8_planes(&mapAhi, T8, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_TILE, I8_NS - I8_A_LO) ||
      !tma::encode_u8_planes(&mapB, T8, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_RB, I8_NS))
    return DISTBO_ECUDA;
  const int sms = device_sm_count();
  I8Args a;
  a.n_rb = n_pad / I8_RB; a.ntiles = 0; a.slices = 1; a.tiles_in_flight = 1;
  {
    const char* dbg = getenv("DISTBO_I8_DEBUG");
    a.debug = dbg ? atoi(dbg) : 0;
  }
  a.rowscale = colscale; a.kexp = nullptr; a.V = nullptr; a.partial = nullptr;
  a.item_off = reinterpret_cast<const int*>(sched);
  a.items = reinterpret_cast<const int4*>(sched + (sms + 1 + 3) / 4 * 4);
  a.C = Kinv; a.ldc = n_pad; a.rowscale_a = colscale; a.alpha = 1.0; a.beta = 0;
  a.meet = nullptr; a.meet_stride = 0; a.meet_waves = 0;
  score_i8_kernel<1><<<sms, I8_THREADS, I8_SMEM, st>>>(mapA, mapAhi, mapB, a);
  DB_CHECK_LAUNCH();
  return DISTBO_OK;
}

// ---- Top levels of the triangular inverse on the tcgen05 tensor cores ------------------------------------------
// Level h of the halving tree merges pairs of inverted diagonal blocks:  W21 = -W22 (L21 W11).  Both products run
// on the digit-plane kernel: T = L21 W11 from the row planes of L and the column planes of W11, then W21 = -W22 T
// from the row planes of W22 and the column planes of T.  Only the levels whose blocks have at least `min_rows`
// rows take this route (the GEMMs of the small levels are launch-bound; they stay on the DMMA tile GEMM).
namespace db {

struct TriI8Level {
  std::vector<int2> cr, rr;            // per column / per row ranges (see i8_col_range, i8_wsplit_kernel)
  std::vector<int4> items[2];          // work items of the two products, per CTA
  std::vector<int> off[2];
};

static void tri_i8_deal(const std::vector<int4>& all, int sms, std::vector<int4>* items, std::vector<int>* off) {
  std::vector<int4> sorted(all);
  std::stable_sort(sorted.begin(), sorted.end(),
                   [](const int4& a, const int4& b) { return (a.w - a.z) > (b.w - b.z); });
  std::vector<double> load(sms, 0.0);
  std::vector<std::vector<int4>> lists(sms);
  for (const int4& it : sorted) {
    int best = 0;
    for (int c = 1; c < sms; ++c)
      if (load[c] < load[best]) best = c;
    load[best] += (double)(it.w - it.z) + 2.0;
    lists[best].push_back(it);
  }
  items->clear();
  off->assign(sms + 1, 0);
  for (int c = 0; c < sms; ++c) {
    (*off)[c] = (int)items->size();
    items->insert(items->end(), lists[c].begin(), lists[c].end());
  }
  (*off)[sms] = (int)items->size();
}

static int tri_i8_levels_build(int n_pad, int min_rows, std::vector<TriI8Level>* out);

// cached per (SM count, n_pad, min_rows): the lists are rebuilt on the host only once (a training loop that is not
// replayed as a CUDA graph enqueues the inverse every epoch)
static int tri_i8_levels(int n_pad, int min_rows, const std::vector<TriI8Level>** out) {
  if (!out) return tri_i8_levels_build(n_pad, min_rows, nullptr);
  static std::mutex mu;
  static std::map<std::tuple<int, int, int>, std::pair<int, std::vector<TriI8Level>>> cache;   // entries never move
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_tuple(device_sm_count(), n_pad, min_rows);
  auto it = cache.find(key);
  if (it == cache.end()) {
    std::vector<TriI8Level> lv;
    const int first = tri_i8_levels_build(n_pad, min_rows, &lv);
    it = cache.emplace(key, std::make_pair(first, std::move(lv))).first;
  }
  *out = &it->second.second;
  return it->second.first;
}

// returns the first height that takes the int8 route (levels.size() when none does) and fills `out`
static int tri_i8_levels_build(int n_pad, int min_rows, std::vector<TriI8Level>* out) {
  const int nb = n_pad / 128, sms = device_sm_count();
  std::vector<std::vector<TriNode>> levels;
  build_tri(0, nb, levels);
  int first = (int)levels.size();
  for (int h = 1; h < (int)levels.size(); ++h) {
    int small = 1 << 30;
    for (const TriNode& nd : levels[h]) small = std::min(small, std::min(nd.n1, nd.n2) * 128);
    if (small >= min_rows) {
      first = h;
      break;
    }
  }
  if (first < 1) first = 1;
  if (out) {
    out->clear();
    for (int h = first; h < (int)levels.size(); ++h) {
      TriI8Level lv;
      lv.cr.assign(n_pad, make_int2(0, 0));
      lv.rr.assign(n_pad, make_int2(0, 0));
      std::vector<int4> g1, g2;
      for (const TriNode& nd : levels[h]) {
        const int o = nd.o * 128, mid = (nd.o + nd.n1) * 128, end = (nd.o + nd.n1 + nd.n2) * 128;
        for (int j = o; j < mid; ++j) lv.cr[j] = make_int2(mid, end);
        for (int i = mid; i < end; ++i) lv.rr[i] = make_int2(mid, end);
        for (int a = mid; a < end; a += I8_TILE)
          for (int b = o; b < mid; b += I8_RB) {
            g1.push_back(make_int4(a, b, b / I8_KC, mid / I8_KC));                 // T = L21 W11 : k in [b, mid)
            g2.push_back(make_int4(a, b, mid / I8_KC, (a + I8_TILE) / I8_KC));     // W21 = -W22 T: k in [mid, a+128)
          }
      }
      tri_i8_deal(g1, sms, &lv.items[0], &lv.off[0]);
      tri_i8_deal(g2, sms, &lv.items[1], &lv.off[1]);
      out->push_back(std::move(lv));
    }
  }
  return first;
}

static size_t pad4(size_t w) { return (w + 3) / 4 * 4; }

}  // namespace db

extern "C" int distbo_trtri_i8_first_level(int n_pad, int min_rows) {
  if (n_pad <= 0 || n_pad % 128 || min_rows < 128) return DISTBO_EARG;
  return tri_i8_levels(n_pad, min_rows, nullptr);
}

extern "C" int64_t distbo_trtri_i8_sched_words(int n_pad, int min_rows) {
  if (n_pad <= 0 || n_pad % 128 || min_rows < 128) return 0;
  const std::vector<TriI8Level>* lvp = nullptr;
  tri_i8_levels(n_pad, min_rows, &lvp);
  const std::vector<TriI8Level>& lv = *lvp;
  size_t w = 4;
  for (const TriI8Level& l : lv)
    w += 4 * (size_t)n_pad + pad4(l.off[0].size()) + 4 * l.items[0].size() + pad4(l.off[1].size()) + 4 * l.items[1].size();
  return (int64_t)w;
}

extern "C" int distbo_trtri_i8_plan(int n_pad, int min_rows, int32_t* sched) {
  if (!sched || n_pad <= 0 || n_pad % 128 || n_pad / I8_RB > I8_MAX_RB || min_rows < 128) return DISTBO_EARG;
  const std::vector<TriI8Level>* lvp = nullptr;
  tri_i8_levels(n_pad, min_rows, &lvp);
  const std::vector<TriI8Level>& lv = *lvp;
  std::vector<int32_t> host;
  auto put = [&](const void* src, size_t words, size_t padded) {
    const size_t at = host.size();
    host.resize(at + padded, 0);
    memcpy(host.data() + at, src, words * sizeof(int32_t));
  };
  for (const TriI8Level& l : lv) {
    put(l.cr.data(), 2 * (size_t)n_pad, 2 * (size_t)n_pad);
    put(l.rr.data(), 2 * (size_t)n_pad, 2 * (size_t)n_pad);
    for (int g = 0; g < 2; ++g) {
      put(l.off[g].data(), l.off[g].size(), pad4(l.off[g].size()));
      put(l.items[g].data(), 4 * l.items[g].size(), 4 * l.items[g].size());
    }
  }
  if (!host.empty()) DB_CUDA(cudaMemcpy(sched, host.data(), host.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  return DISTBO_OK;
}

extern "C" int64_t distbo_trtri_i8_workspace(int n_pad) {
  // doubles: four sets of digit planes (rows of L, columns of W11, rows of W22, columns of T) + their exponents
  return 4 * (int64_t)n_pad * n_pad + 4 * (int64_t)n_pad;
}

extern "C" int distbo_trtri_i8_f64(const double* L, double* W, double* T, int n_pad, int min_rows,
                                   const int32_t* sched, double* ws, void* stream) {
  if (!L || !W || !T || !sched || !ws || n_pad <= 0 || n_pad % 128 || n_pad / I8_RB > I8_MAX_RB || min_rows < 128)
    return DISTBO_EARG;
  cudaStream_t st = (cudaStream_t)stream;
  static PerDeviceOnce once;
  if (once.needed()) {
    DB_CUDA(cudaFuncSetAttribute(score_i8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM));
    once.mark();
  }
  const std::vector<TriI8Level>* lvp = nullptr;
  tri_i8_levels(n_pad, min_rows, &lvp);
  const std::vector<TriI8Level>& lv = *lvp;
  if (lv.empty()) return DISTBO_OK;
  const size_t nn = (size_t)n_pad * n_pad;
  signed char* L8 = reinterpret_cast<signed char*>(ws);
  signed char* C8 = reinterpret_cast<signed char*>(ws + nn);        // columns of W11
  signed char* R8 = reinterpret_cast<signed char*>(ws + 2 * nn);    // rows of W22
  signed char* T8 = reinterpret_cast<signed char*>(ws + 3 * nn);    // columns of T
  double* rsL = ws + 4 * nn;
  double* csW = rsL + n_pad;
  double* rsW = csW + n_pad;
  double* csT = rsW + n_pad;
  CUtensorMap mL, mLhi, mC, mR, mRhi, mT;
  if (!tma::encode_u8_planes(&mL, L8, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_TILE, I8_A_LO) ||
      !tma::encode_u8_planes(&mLhi, L8, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_TILE, I8_NS - I8_A_LO) ||
      !tma::encode_u8_planes(&mC, C8, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_RB, I8_NS) ||
      !tma::encode_u8_planes(&mR, R8, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_TILE, I8_A_LO) ||
      !tma::encode_u8_planes(&mRhi, R8, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_TILE, I8_NS - I8_A_LO) ||
      !tma::encode_u8_planes(&mT, T8, (uint64_t)n_pad, (uint64_t)n_pad, I8_NS, I8_RB, I8_NS))
    return DISTBO_ECUDA;
  const int sms = device_sm_count();
  const dim3 cgrid(n_pad / 64, (n_pad + 511) / 512), tgrid(n_pad / 64, n_pad / 64);
  // The row splits (of L once, of the W22 blocks per level) do not depend on the product that runs before them on the
  // main stream: they go to a side stream and are joined where their planes are first read (fork / join by events,
  // capture-safe; stream and events live for this call, their destruction is deferred by the runtime).
  struct Side {
    cudaStream_t s = nullptr;
    std::vector<cudaEvent_t> ev;
    ~Side() {
      for (cudaEvent_t e : ev) cudaEventDestroy(e);
      if (s) cudaStreamDestroy(s);
    }
    cudaEvent_t event() {
      cudaEvent_t e = nullptr;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess) ev.push_back(e);
      return e;
    }
  } side;
  DB_CUDA(cudaStreamCreateWithFlags(&side.s, cudaStreamNonBlocking));
  cudaEvent_t ev_fork = side.event(), ev_l = side.event();
  if (!ev_fork || !ev_l) return DISTBO_ECUDA;
  DB_CUDA(cudaEventRecord(ev_fork, st));
  DB_CUDA(cudaStreamWaitEvent(side.s, ev_fork, 0));
  // rows of L, once (k <= i)
  i8_wsplit_kernel<<<n_pad, 256, 0, side.s>>>(L, n_pad, L8, rsL, nullptr);
  DB_CHECK_LAUNCH();
  DB_CUDA(cudaEventRecord(ev_l, side.s));
  bool l_joined = false;
  const int32_t* cur = sched;
  for (const TriI8Level& l : lv) {
    const int2* cr = reinterpret_cast<const int2*>(cur);
    const int2* rr = reinterpret_cast<const int2*>(cur + 2 * (size_t)n_pad);
    cur += 4 * (size_t)n_pad;
    const int* off1 = cur;
    cur += pad4(l.off[0].size());
    const int4* it1 = reinterpret_cast<const int4*>(cur);
    cur += 4 * l.items[0].size();
    const int* off2 = cur;
    cur += pad4(l.off[1].size());
    const int4* it2 = reinterpret_cast<const int4*>(cur);
    cur += 4 * l.items[1].size();

    I8Args a;
    a.n_rb = n_pad / I8_RB; a.ntiles = 0; a.slices = 1; a.tiles_in_flight = 1; a.debug = 0;
    a.kexp = nullptr; a.V = nullptr; a.partial = nullptr;
    a.meet = nullptr; a.meet_stride = 0; a.meet_waves = 0;
    a.ldc = n_pad; a.beta = 0;
    // columns of W11, then T = L21 W11
    DB_CUDA(cudaMemsetAsync(csW, 0, sizeof(double) * n_pad, st));
    i8_colmax_kernel<<<cgrid, 256, 0, st>>>(W, n_pad, reinterpret_cast<unsigned long long*>(csW), cr, 0);
    DB_CHECK_LAUNCH();
    i8_colscale_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(csW, n_pad);
    DB_CHECK_LAUNCH();
    i8_wtsplit_kernel<<<tgrid, 256, 0, st>>>(W, n_pad, csW, C8, cr, 0);
    DB_CHECK_LAUNCH();
    // rows of W22 on the side stream, next to the first product (W of the previous level is complete: the side
    // stream waits for the point the main stream has reached, i.e. behind the previous level's second product)
    cudaEvent_t ev_lvl = side.event(), ev_r = side.event();
    if (!ev_lvl || !ev_r) return DISTBO_ECUDA;
    DB_CUDA(cudaEventRecord(ev_lvl, st));
    DB_CUDA(cudaStreamWaitEvent(side.s, ev_lvl, 0));
    i8_wsplit_kernel<<<n_pad, 256, 0, side.s>>>(W, n_pad, R8, rsW, rr);
    DB_CHECK_LAUNCH();
    DB_CUDA(cudaEventRecord(ev_r, side.s));
    if (!l_joined) {
      DB_CUDA(cudaStreamWaitEvent(st, ev_l, 0));
      l_joined = true;
    }
    a.item_off = off1; a.items = it1; a.C = T; a.alpha = 1.0; a.rowscale_a = rsL; a.rowscale = csW;
    score_i8_kernel<1><<<sms, I8_THREADS, I8_SMEM, st>>>(mL, mLhi, mC, a);
    DB_CHECK_LAUNCH();
    // columns of T, then W21 = -W22 T
    DB_CUDA(cudaMemsetAsync(csT, 0, sizeof(double) * n_pad, st));
    i8_colmax_kernel<<<cgrid, 256, 0, st>>>(T, n_pad, reinterpret_cast<unsigned long long*>(csT), cr, 1);
    DB_CHECK_LAUNCH();
    i8_colscale_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(csT, n_pad);
    DB_CHECK_LAUNCH();
    i8_wtsplit_kernel<<<tgrid, 256, 0, st>>>(T, n_pad, csT, T8, cr, 1);
    DB_CHECK_LAUNCH();
    DB_CUDA(cudaStreamWaitEvent(st, ev_r, 0));
    a.item_off = off2; a.items = it2; a.C = W; a.alpha = -1.0; a.rowscale_a = rsW; a.rowscale = csT;
    score_i8_kernel<1><<<sms, I8_THREADS, I8_SMEM, st>>>(mR, mRhi, mT, a);
    DB_CHECK_LAUNCH();
  }
  return DISTBO_OK;
}
