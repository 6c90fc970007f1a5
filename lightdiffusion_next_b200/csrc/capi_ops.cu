// C ABI: error plumbing + op-level entry points (include/ldn.h). No torch types cross this boundary.
#include "../../include/ldn.h"
#include "common.h"

#include <mutex>

namespace ldn {
static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
}  // namespace ldn

#define LDN_API_BEGIN try {
#define LDN_API_END                     \
  }                                     \
  catch (const std::exception& e) {     \
    ldn::set_last_error(e.what());      \
    return 1;                           \
  }                                     \
  catch (...) {                         \
    ldn::set_last_error("unknown error"); \
    return 2;                           \
  }                                     \
  return 0;

using namespace ldn;

static const size_t kOpWsBytes = (size_t)64 << 20;
static float* op_splitk_ws() {  // workspace for op-level calls (the engine has its own per program)
  static float* ws = nullptr;
  if (!ws) LDN_CUDA(cudaMalloc(&ws, kOpWsBytes));
  return ws;
}

extern "C" {

const char* ldn_last_error(void) { return g_last_error.c_str(); }
int ldn_version(void) { return 100; }

int ldn_gemm_bf16(const void* A0, int64_t lda0, int K0, const void* A1, int64_t lda1, int K1, const void* Wt, int M,
                  int N, const float* bias, const float* rowbias, int ld_rowbias, int rows_per_batch,
                  const void* residual, int64_t ldr, void* out, int64_t ldo, float* out_f32, int epi, int head_dim,
                  int head_slot, int BN, void* stream) {
  LDN_API_BEGIN
  GemmArgs a;
  a.A0 = (const bf16*)A0; a.lda0 = lda0; a.K0 = K0;
  a.A1 = (const bf16*)A1; a.lda1 = lda1; a.K1 = K1;
  a.Wt = (const bf16*)Wt; a.M = M; a.N = N;
  a.bias = bias; a.rowbias = rowbias; a.ld_rowbias = ld_rowbias; a.rows_per_batch = rows_per_batch;
  a.residual = (const bf16*)residual; a.ldr = ldr;
  a.out = (bf16*)out; a.ldo = ldo; a.out_f32 = out_f32;
  a.epi = epi; a.head_dim = head_dim; a.head_slot = head_slot; a.BN = BN;
  a.splitk_ws = op_splitk_ws(); a.splitk_ws_bytes = kOpWsBytes;
  GemmPlan plan = make_gemm_plan(a);
  launch_gemm(plan, (cudaStream_t)stream);
  LDN_API_END
}

int ldn_conv3x3_bf16(const void* x, const void* Wt, int B, int H, int W, int Cin, int Cout, const float* bias,
                     const float* rowbias, int ld_rowbias, const void* residual, void* out, void* stream) {
  LDN_API_BEGIN
  GemmArgs a;
  a.conv = true;
  a.A0 = (const bf16*)x; a.B = B; a.H = H; a.W = W; a.Cin = Cin;
  a.Wt = (const bf16*)Wt; a.N = Cout; a.M = B * H * W;
  a.bias = bias; a.rowbias = rowbias; a.ld_rowbias = ld_rowbias;
  a.residual = (const bf16*)residual; a.ldr = Cout;
  a.out = (bf16*)out; a.ldo = Cout;
  a.splitk_ws = op_splitk_ws(); a.splitk_ws_bytes = kOpWsBytes;
  GemmPlan plan = make_gemm_plan(a);
  launch_gemm(plan, (cudaStream_t)stream);
  LDN_API_END
}

}  // extern "C"

extern "C" int ldn_attention_bf16(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* Vt, int64_t ldvt,
                                  int64_t vt_rows, int vt_head_stride, int B, int heads, int Nq, int Nk, int nk_pad,
                                  int d, int slot, int causal, float scale, void* out, int64_t ldo, void* stream) {
  LDN_API_BEGIN
  AttnArgs a;
  a.Q = (const bf16*)Q; a.ldq = ldq; a.K = (const bf16*)K; a.ldk = ldk; a.Vt = (const bf16*)Vt; a.ldvt = ldvt;
  a.vt_rows = vt_rows; a.vt_head_stride = vt_head_stride; a.B = B; a.heads = heads; a.Nq = Nq; a.Nk = Nk; a.nk_pad = nk_pad; a.d = d; a.slot = slot;
  a.causal = causal & 1; a.fold = (causal & 2) ? 1 : 0; a.scale = scale; a.out = (bf16*)out; a.ldo = ldo;
  AttnPlan plan = make_attn_plan(a);
  launch_attn(plan, (cudaStream_t)stream);
  LDN_API_END
}
