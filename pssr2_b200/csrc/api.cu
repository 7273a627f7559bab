// C ABI plumbing: error strings, launch counter, plan create / run / destroy.
#include <stdarg.h>
#include <string.h>
#include <new>
#include "common.cuh"
#include "plan.h"

namespace pssr {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int device_sm_count() {
  static int sms[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    if (cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms[dev] <= 0) sms[dev] = 148;
  }
  return sms[dev];
}

}  // namespace pssr

using namespace pssr;

extern "C" {

const char* pssr_last_error(void) { return g_err; }
const char* pssr_version(void) { return "pssr_b200 0.1 sm_100a"; }
int64_t pssr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int pssr_debug_trace(int64_t* out, int64_t n) {
  PSSR_REQUIRE(out != nullptr && n > 0, PSSR_EINVAL, "debug_trace: bad arguments");
  return v3_trace_fetch(reinterpret_cast<long long*>(out), (int)n);
}

int pssr_plan_create(const pssr_op_t* ops, int32_t n_ops, int32_t dtype, pssr_plan_t** out) {
  PSSR_REQUIRE(ops != nullptr && n_ops > 0 && out != nullptr, PSSR_EINVAL, "plan_create: bad arguments");
  PSSR_REQUIRE(dtype == PSSR_DT_BF16 || dtype == PSSR_DT_FP16, PSSR_EINVAL, "plan_create: dtype must be PSSR_DT_BF16/FP16");
  int cc_major = 0, dev = 0;
  PSSR_CHECK_CUDA(cudaGetDevice(&dev));
  PSSR_CHECK_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  PSSR_REQUIRE(cc_major == 10, PSSR_EUNSUP, "plan_create: this library is built for sm_100a only (device is sm_%d0)", cc_major);
  pssr_plan* plan = new (std::nothrow) pssr_plan();
  PSSR_REQUIRE(plan != nullptr, PSSR_EINVAL, "plan_create: out of host memory");
  plan->dtype = dtype;
  plan->ops.assign(ops, ops + n_ops);
  plan->conv_index.assign(n_ops, -1);
  for (int i = 0; i < n_ops; ++i) {
    const pssr_op_t& op = plan->ops[i];
    switch (op.kind) {
      case PSSR_OP_CONV: {
        ConvOp c;
        int rc = v3_supported(op.u.conv) ? v3_prepare(op.u.conv, dtype, c) : conv_prepare(op.u.conv, dtype, c);
        if (rc != PSSR_OK) {
          char msg[400];
          snprintf(msg, sizeof(msg), "%s", g_err);
          set_error("op %d: %s", i, msg);
          delete plan;
          return rc;
        }
        plan->conv_index[i] = (int)plan->convs.size();
        plan->convs.push_back(c);
        break;
      }
      case PSSR_OP_PREP:
      case PSSR_OP_MAXPOOL:
      case PSSR_OP_TAIL:
      case PSSR_OP_TAILSUM:
      case PSSR_OP_STEM:
      case PSSR_OP_LAYERNORM:
      case PSSR_OP_DWCONV_LN:
      case PSSR_OP_ESE:
      case PSSR_OP_CAST8:
      case PSSR_OP_RESAMPLE:
      case PSSR_OP_WINATTN:
        break;
      default:
        set_error("plan_create: op %d has unsupported kind %d", i, op.kind);
        delete plan;
        return PSSR_EUNSUP;
    }
  }
  if (!plan->convs.empty()) {
    const size_t bytes = plan->convs.size() * kTmapsPerConv * sizeof(CUtensorMap);
    cudaError_t e = cudaMalloc(&plan->tmaps_dev, bytes);
    if (e != cudaSuccess) {
      set_error("plan_create: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
      delete plan;
      return PSSR_ECUDA;
    }
    for (size_t i = 0; i < plan->convs.size(); ++i) {
      e = cudaMemcpy(reinterpret_cast<uint8_t*>(plan->tmaps_dev) + i * kTmapsPerConv * sizeof(CUtensorMap), plan->convs[i].tmaps,
                     kTmapsPerConv * sizeof(CUtensorMap), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) {
        set_error("plan_create: tensor-map upload failed: %s", cudaGetErrorString(e));
        cudaFree(plan->tmaps_dev);
        delete plan;
        return PSSR_ECUDA;
      }
    }
  }
  *out = plan;
  return PSSR_OK;
}

int pssr_plan_run_range(pssr_plan_t* plan, int32_t first, int32_t count, void* stream) {
  PSSR_REQUIRE(plan != nullptr, PSSR_EINVAL, "plan_run: null plan");
  PSSR_REQUIRE(first >= 0 && count >= 0 && first + count <= (int)plan->ops.size(), PSSR_EINVAL, "plan_run: op range out of bounds");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int i = first; i < first + count; ++i) {
    const pssr_op_t& op = plan->ops[i];
    int rc = PSSR_OK;
    switch (op.kind) {
      case PSSR_OP_CONV: {
        const int ci = plan->conv_index[i];
        const void* tm = reinterpret_cast<uint8_t*>(plan->tmaps_dev) + (size_t)ci * kTmapsPerConv * sizeof(CUtensorMap);
        const int variant = plan->convs[ci].variant;
        rc = variant == 3 ? v3_launch(plan->convs[ci], tm, st) : conv_launch(plan->convs[ci], tm, st);
        break;
      }
      case PSSR_OP_PREP:
        rc = prep_launch(op.u.prep, plan->dtype, st);
        break;
      case PSSR_OP_MAXPOOL:
        rc = pool_launch(op.u.pool, plan->dtype, st);
        break;
      case PSSR_OP_TAIL:
        rc = tail_launch(op.u.tail, plan->dtype, st);
        break;
      case PSSR_OP_TAILSUM:
        rc = tailsum_launch(op.u.tailsum, st);
        break;
      case PSSR_OP_STEM:
        rc = stem_launch(op.u.stem, plan->dtype, st);
        break;
      case PSSR_OP_LAYERNORM:
        rc = ln_launch(op.u.ln, plan->dtype, st);
        break;
      case PSSR_OP_DWCONV_LN:
        rc = dwln_launch(op.u.dwln, plan->dtype, st);
        break;
      case PSSR_OP_ESE:
        rc = ese_launch(op.u.ese, plan->dtype, st);
        break;
      case PSSR_OP_CAST8:
        rc = cast8_launch(op.u.cast8, plan->dtype, st);
        break;
      case PSSR_OP_RESAMPLE:
        rc = resample_launch(op.u.resample, plan->dtype, st);
        break;
      case PSSR_OP_WINATTN:
        rc = winattn_launch(op.u.winattn, plan->dtype, st);
        break;
      default:
        set_error("plan_run: op %d has unsupported kind %d", i, op.kind);
        rc = PSSR_EUNSUP;
    }
    if (rc != PSSR_OK) return rc;
  }
  return PSSR_OK;
}

int pssr_plan_run(pssr_plan_t* plan, void* stream) {
  PSSR_REQUIRE(plan != nullptr, PSSR_EINVAL, "plan_run: null plan");
  return pssr_plan_run_range(plan, 0, (int)plan->ops.size(), stream);
}

int32_t pssr_plan_num_ops(const pssr_plan_t* plan) { return plan ? (int32_t)plan->ops.size() : 0; }

void pssr_plan_destroy(pssr_plan_t* plan) {
  if (!plan) return;
  if (plan->tmaps_dev) cudaFree(plan->tmaps_dev);
  delete plan;
}

}  // extern "C"
