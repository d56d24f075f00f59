// Host-side helpers shared by the C-ABI translation units: error text, argument checks, TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <unordered_map>

#include "../../include/anyv2v_b200.h"

namespace av2v {

char* last_error_buf();  // thread-local 512-byte buffer (defined in abi.cu)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define AV2V_CHECK_CUDA(expr)                                                                   \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return ::av2v::fail(AV2V_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                          __FILE__, __LINE__);                                                  \
  } while (0)

#define AV2V_REQUIRE(cond, code, ...) \
  do {                                \
    if (!(cond)) return ::av2v::fail(code, __VA_ARGS__); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency, so the
// library also loads on a box without a driver — needed for the CPU-side symbol-export test).
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// fp16 tiled tensor map with 128-byte swizzle; dims/box innermost first; strides (bytes) for dims 1..rank-1.
inline int make_tmap_f16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                         const uint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return fail(AV2V_ECUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_b[i];
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                   gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(AV2V_ECUDA,
                "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u] base %p",
                (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
  }
  return AV2V_OK;
}

int sm_count_cached();  // abi.cu

// Round-2 candidates are selected per call by environment switches (read at call time so one process can A/B them);
// unset = the shipped, GPU-verified path.
inline int env_int(const char* name, int dflt = 0) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
// AV2V_PDL=1: launch with programmatic stream serialisation (the kernels then execute griddepcontrol.wait before
// their first global access, so barrier init / TMEM allocation / descriptor prefetch overlap the predecessor's tail)
inline int pdl_enabled() { return env_int("AV2V_PDL") ? 1 : 0; }

// AV2V_PINGPONG=1 (round-2 candidate): choose every launch's traversal direction OPPOSITE to the direction in which its input
// was last written.  At B = 3 the activations of the 64 x 64 level are 126 MB — the size of L2.  Every kernel walks its tiles /
// rows front to back, so under an LRU-like policy a consumer misses on the head of the tensor its producer just wrote and evicts
// the tail before reaching it.  A consumer that walks back to front hits on the resident tail — and leaves, in turn, the head of
// its own output for a successor that walks forward.  Directions only permute the order of independent tiles / rows: results are
// unchanged.  The library remembers, per output pointer, the direction of the last write by one of its kernels; an unknown
// producer (a torch op, a copy) is assumed to have written front to back.  `dflt` = direction without the switch (0 = forward
// everywhere on the shipped path).  Decisions made during CUDA-graph capture are baked into the graph.
inline std::unordered_map<const void*, int>& direction_table() {
  static std::unordered_map<const void*, int> table;
  return table;
}
inline void record_direction(const void* out, int dir) {
  if (out == nullptr || env_int("AV2V_PINGPONG") != 1) return;
  auto& t = direction_table();
  if (t.size() > 16384) t.clear();
  t[out] = dir;
}
inline int pick_direction(const void* in, const void* out, int dflt = 0) {
  if (env_int("AV2V_PINGPONG") != 1) return dflt;
  auto& t = direction_table();
  const auto it = t.find(in);
  const int dir = (it == t.end()) ? 1 : !it->second;  // unknown producer: assume it wrote front to back -> read back to front
  record_direction(out, dir);
  return dir;
}

// Kernel launch with optional programmatic stream serialisation (PDL) and an optional cluster of `cluster_x` CTAs.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int pdl,
                             int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = static_cast<unsigned>(cluster_x);
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = static_cast<unsigned>(na);
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// attention2q_tcgen05.cu: two-query-tile attention (rows mode, n_v = 1), AV2V_ATTN_2Q = 1 | 2 | 3
int attn2q_launch(const ::av2v_attn_args* a, int mode, int pdl, cudaStream_t stream);
// attention_v10_tcgen05.cu: v9 with P in its own TMEM columns and early S issue (all modes, n_v = 1 | 3), AV2V_ATTN_V10 = 1
int attn_v10_launch(const ::av2v_attn_args* a, int mode, int pdl, cudaStream_t stream);  // mode 2: + FMA-pipe exp2 (25 %, scalar); 3: packed fp32x2 + 3/8

}  // namespace av2v
