// C-ABI glue: version, thread-local error string, argument validation shared by the kernels.
#include <stdarg.h>

#include "qasr_common.cuh"

namespace {
thread_local char g_err[512] = "";
// bring-up / benchmarking knobs (qasr_debug_set; not part of the public header): -1 = unset
//   0,1: B-operand descriptor strides   2,3: TMA coordinate experiments   4: force streaming weights
//   5: M tiles per CTA   6: skip epilogue (1) / MMA (2) / stores (4)   8: depthwise path (0 = IMMA, 1 = tcgen05)
//   16: IMMA depthwise rows end with the masked tail path (1) instead of the shifted full tile
constexpr int kKnobs = 24;
long long g_dbg[kKnobs] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
}

extern "C" int qasr_debug_set(int key, long long value) {
  if (key < 0 || key >= kKnobs) return QASR_ERR_INVALID;
  g_dbg[key] = value;
  return QASR_OK;
}
long long qasr_debug_get(int key) { return key >= 0 && key < kKnobs ? g_dbg[key] : -1; }
// how many bring-up knobs are set: bench.py refuses to time the kernels unless this is 0
extern "C" int qasr_debug_knobs_set(void) {
  int n = 0;
  for (int i = 0; i < kKnobs; ++i) n += g_dbg[i] != -1;
  return n;
}

namespace { thread_local bool g_plain_launch = false; }
void qasr_set_plain_launch(bool plain) { g_plain_launch = plain; }
bool qasr_plain_launch() { return g_plain_launch; }

namespace {
unsigned long long* g_prof_buf = nullptr;
int g_prof_cap = 0, g_prof_n = 0;
}
unsigned long long* qasr_next_stamp() {
  if (g_prof_buf == nullptr || g_prof_n >= g_prof_cap) return nullptr;
  return g_prof_buf + 2 * (g_prof_n++);
}
// stamps: device buffer of `capacity` (start, end) uint64 pairs; the caller resets every pair to (UINT64_MAX >> 1, 0)
// before each run.  Launches issued (or captured into a CUDA graph) between begin and end carry their slot with them.
extern "C" int qasr_profile_begin(void* stamps, int capacity) {
  QASR_CHECK_ARG(stamps != nullptr && capacity > 0, "profile_begin: bad arguments");
  g_prof_buf = static_cast<unsigned long long*>(stamps);
  g_prof_cap = capacity;
  g_prof_n = 0;
  return QASR_OK;
}
extern "C" int qasr_profile_end(void) {
  const int n = g_prof_n;
  g_prof_buf = nullptr;
  g_prof_cap = g_prof_n = 0;
  return n;
}

void qasr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int qasr_version(void) { return QASR_VERSION; }
extern "C" const char* qasr_last_error(void) { return g_err; }

extern "C" int qasr_sm_count(void) {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int qasr_validate_epilogue(const qasr_epilogue* e, bool allow_second_operand) {
  QASR_CHECK_ARG(e->mode >= QASR_EPI_REQUANT_I8 && e->mode <= QASR_EPI_STATS, "epilogue: unknown mode %d", e->mode);
  QASR_CHECK_ARG(e->ch0 != nullptr && qasr_aligned16(e->ch0), "epilogue: ch0 missing or not 16-byte aligned");
  QASR_CHECK_ARG(allow_second_operand || e->ch1 == nullptr, "epilogue: this kernel takes one conv operand");
  QASR_CHECK_ARG(e->ch1 == nullptr || qasr_aligned16(e->ch1), "epilogue: ch1 not 16-byte aligned");
  if (e->mode == QASR_EPI_REQUANT_I8) {
    QASR_CHECK_ARG(e->qmin < e->qmax && e->qmin >= -256 && e->qmax <= 255, "epilogue: clamp [%d, %d]", e->qmin, e->qmax);
    const bool s2 = e->flags & QASR_F_STAGE2;
    // codes are non-negative after a ReLU (the reference's "act_bit + 1" unsigned inputs,
    // jasper.py:159-164): a [-256, 255] clamp then means an unsigned 8-bit tensor
    const bool nonneg = (e->flags & QASR_F_RELU1) ||
                        ((e->flags & QASR_F_RELU0) && e->ch1 == nullptr && e->prev == nullptr);
    int lo = s2 ? e->qmin2 : e->qmin;
    if (nonneg && lo < 0) lo = 0;
    const int hi = s2 ? e->qmax2 : e->qmax;
    QASR_CHECK_ARG(hi - lo <= 255 && lo >= -128 && hi <= 255 && !(lo < 0 && hi > 127),
                   "epilogue: final code range [%d, %d] does not fit one 8-bit tensor", lo, hi);
    QASR_CHECK_ARG(!s2 || (e->e2 >= 1 && e->e2 <= 62 && e->m2 > 0), "epilogue: stage-2 dyadic (m2=%d, e2=%d)", e->m2, e->e2);
    QASR_CHECK_ARG(!(e->flags & QASR_F_MASK) || e->out_lens != nullptr, "epilogue: MASK needs out_lens");
    QASR_CHECK_ARG(e->prev == nullptr || (e->prev_ld % 16 == 0 && qasr_aligned16(e->prev)), "epilogue: prev pitch/alignment");
  } else if (e->mode == QASR_EPI_STATS) {
    QASR_CHECK_ARG(e->prev == nullptr, "epilogue: STATS takes conv operands only");
    QASR_CHECK_ARG(!(e->flags & QASR_F_MASK) || e->out_lens != nullptr, "epilogue: MASK needs out_lens");
  } else {
    QASR_CHECK_ARG(e->ch1 == nullptr && e->prev == nullptr, "epilogue: residual operands need REQUANT mode");
  }
  QASR_CHECK_ARG(e->f32_mult == nullptr ||
                     (e->mode == QASR_EPI_REQUANT_I8 && e->ch1 == nullptr && e->prev == nullptr &&
                      !(e->flags & (QASR_F_STAGE2 | QASR_F_EXACT_F32 | QASR_F_EXACT_DYAD | QASR_F_BIAS64))),
                 "epilogue: f32_mult belongs to the single-operand, single-stage integer fast path");
  QASR_CHECK_ARG(!(e->flags & QASR_F_FAST_F32) ||
                     (e->mode == QASR_EPI_REQUANT_I8 && e->ch1 == nullptr && e->prev == nullptr &&
                      !(e->flags & (QASR_F_STAGE2 | QASR_F_EXACT_F32 | QASR_F_EXACT_DYAD | QASR_F_BIAS64))),
                 "epilogue: FAST_F32 is a single-operand, single-stage REQUANT fast path");
  QASR_CHECK_ARG(!(e->flags & QASR_F_RANGE_CHECK) ||
                     (e->mode == QASR_EPI_REQUANT_I8 && e->f32_mult == nullptr &&
                      !(e->flags & (QASR_F_FAST_F32 | QASR_F_EXACT_F32 | QASR_F_EXACT_DYAD | QASR_F_BIAS64 | QASR_F_DEVICE_TABLES))),
                 "epilogue: RANGE_CHECK belongs to the integer fast path of a REQUANT launch");
  QASR_CHECK_ARG(!(e->flags & QASR_F_CLAMP16) || (e->mode == QASR_EPI_REQUANT_I8 && (e->flags & QASR_F_RELU0) && e->qmin <= 0 &&
                                                  e->ch1 == nullptr && e->prev == nullptr),
                 "epilogue: CLAMP16 belongs to a single-operand REQUANT launch with the ReLU-folded clamp [0, qmax]");
  QASR_CHECK_ARG(!(e->flags & QASR_F_STAGE2_16) ||
                     (e->mode == QASR_EPI_REQUANT_I8 && (e->flags & QASR_F_STAGE2) && (e->flags & QASR_F_RELU1) && e->qmin <= 0 &&
                      e->qmax <= 127 && e->qmin2 <= 0 && e->qmax2 >= 0 && e->e2 >= 17 &&
                      !(e->flags & (QASR_F_EXACT_F32 | QASR_F_EXACT_DYAD | QASR_F_BIAS64))),
                 "epilogue: STAGE2_16 belongs to a fast-path STAGE2 + RELU1 launch with qmax <= 127 and e2 >= 17");
  QASR_CHECK_ARG(!(e->flags & QASR_F_BIAS64) || (e->flags & QASR_F_EXACT_F32) || e->mode != QASR_EPI_REQUANT_I8,
                 "epilogue: BIAS64 requires EXACT_F32");
  return QASR_OK;
}
