// Integer host logic of the path: partial/window arithmetic and per-clip frame counts.
// Follows voice_encoder.py:54-81 (get_num_wins / get_frame_step), melspec.py:50 (T_ve),
// torchaudio kaldi.py:63-67 (snip_edges frame count) and xvector.py:221-231, 364-372.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

#include "cbx_internal.h"

namespace cbx {

int ve_frame_step(double overlap, double rate) {
  if (!(overlap >= 0.0 && overlap < 1.0)) return -1;
  double v = rate > 0.0 ? (double(kSR) / rate) / kVePartial : kVePartial * (1.0 - overlap);
  int step = (int)std::nearbyint(v);   // np.round: ties to even, as nearbyint in the default rounding mode
  if (!(step > 0 && step <= kVePartial)) return -1;
  return step;
}

void ve_num_wins(int64_t n_frames, int step, double min_cov, int64_t* n, int64_t* target) {
  int64_t x = n_frames - kVePartial + step;
  if (x < 0) x = 0;
  int64_t wins = x / step, rem = x % step;
  if (wins == 0 || double(rem + (kVePartial - step)) / double(kVePartial) >= min_cov) wins += 1;
  *n = wins;
  *target = kVePartial + (int64_t)step * (wins - 1);
}

}  // namespace cbx

extern "C" {

int cbx_ve_frame_step(double overlap, double rate) { return cbx::ve_frame_step(overlap, rate); }

int cbx_ve_num_wins(int64_t n_frames, int step, double min_coverage, int64_t* n_wins, int64_t* target_n) {
  if (n_frames <= 0 || step <= 0 || step > cbx::kVePartial || !n_wins || !target_n) return CBX_ERR_ARG;
  cbx::ve_num_wins(n_frames, step, min_coverage, n_wins, target_n);
  return CBX_OK;
}

int cbx_plan_clip(int64_t n, int step, double min_cov, cbx_clip_plan* out) {
  if (!out || n < 0 || step <= 0 || step > cbx::kVePartial) return CBX_ERR_ARG;
  out->n_samples = n;
  out->ve_frames = 1 + n / cbx::kVeHop;
  cbx::ve_num_wins(out->ve_frames, step, min_cov, &out->ve_partials, &out->ve_target);
  out->xv_frames = n < cbx::kKWin ? 0 : 1 + (n - cbx::kKWin) / cbx::kKHop;
  out->xv_tdnn = out->xv_frames > 0 ? (out->xv_frames - 1) / 2 + 1 : 0;
  out->xv_segments = (out->xv_tdnn + cbx::kSegLen - 1) / cbx::kSegLen;
  return CBX_OK;
}

int64_t cbx_trim_num_frames(int64_t n) { return n < 0 ? 0 : 1 + n / cbx::kTrimHop; }

double cbx_clip_cost(int64_t n) {
  cbx_clip_plan p;
  if (cbx_plan_clip(n, 77, 0.8, &p) != CBX_OK) return 0.0;
  // LSTM per partial-step, FCM per frame, TDNN stack per T' frame
  return 2703360.0 * 160.0 * (double)p.ve_partials + 4776960.0 * (double)p.xv_frames + 12959744.0 * (double)p.xv_tdnn;
}

int cbx_partition(const int64_t* n_samples, int64_t n, int world, int32_t* rank_of, int64_t* row_of, double* rank_cost) {
  if (n < 0 || world <= 0 || (n > 0 && (!n_samples || !rank_of))) return CBX_ERR_ARG;
  std::vector<double> cost((size_t)n);
  // most expensive first (ties keep clip order), dealt out and back again over the ranks: rank k receives positions
  // k, 2R-1-k, 2R+k, ... of the sorted list, so counts differ by at most one and every pair of rounds evens the cost out.
  // The order is a stable LSD radix sort on the inverted bit pattern of the (non-negative) cost -- monotone in the value --
  // 1e5 clips: ~2 ms instead of the 15 ms of a comparison sort through an index indirection.
  std::vector<uint64_t> key((size_t)n), key2((size_t)n);
  std::vector<int64_t> order((size_t)n), order2((size_t)n);
  uint64_t differ = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (n_samples[i] < 0) return CBX_ERR_ARG;
    const double cst = cbx_clip_cost(n_samples[i]);
    cost[(size_t)i] = cst;
    uint64_t bits;
    std::memcpy(&bits, &cst, sizeof(bits));
    key[(size_t)i] = ~bits;
    order[(size_t)i] = i;
    differ |= key[(size_t)i] ^ key[0];
  }
  for (int shift = 0; shift < 64; shift += 8) {
    if (((differ >> shift) & 0xff) == 0) continue;          // every key has the same byte here
    size_t hist[257] = {0};
    for (int64_t i = 0; i < n; ++i) ++hist[((key[(size_t)i] >> shift) & 0xff) + 1];
    for (int b = 0; b < 256; ++b) hist[b + 1] += hist[b];
    for (int64_t i = 0; i < n; ++i) {
      const size_t d = hist[(key[(size_t)i] >> shift) & 0xff]++;
      key2[d] = key[(size_t)i]; order2[d] = order[(size_t)i];
    }
    key.swap(key2); order.swap(order2);
  }
  if (rank_cost) std::fill(rank_cost, rank_cost + world, 0.0);
  for (int64_t pos = 0; pos < n; ++pos) {
    const int64_t round = pos / world, k = pos % world;
    const int r = (int)((round & 1) ? world - 1 - k : k);
    rank_of[order[(size_t)pos]] = r;
    if (rank_cost) rank_cost[r] += cost[(size_t)order[(size_t)pos]];
  }
  if (row_of) {   // a rank keeps its clips in ascending clip order
    std::vector<int64_t> count((size_t)world, 0);
    for (int64_t i = 0; i < n; ++i) row_of[i] = count[(size_t)rank_of[i]]++;
  }
  return CBX_OK;
}

const char* cbx_version(void) { return "cbx 0.1 (sm_100a)"; }

}  // extern "C"
