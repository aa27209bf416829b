// distance.cu -- query x gallery distance matrix (torchreid/metrics/distance.py:11-89) on tcgen05.
//
//   euclidean : out[i,j] = (|q_i|^2 + |g_j|^2) - 2 q_i.g_j      (no clamp, no sqrt; distance.py:70-72)
//   cosine    : out[i,j] = 1 - q^_i . g^_j,  x^ = x / max(|x|_2, 1e-12)   (distance.py:86-88)
//
// Pipeline: split.cu turns both operands into bf16 planes (and norms / normalised rows) in one
// pass each, then one persistent tcgen05 GEMM (gemm_sm100.cuh) with the distance epilogue fused.
#include "gemm_sm100.cuh"

namespace agrl {

// One prepared operand: bf16 planes [P][rows][K_pad] followed by the per-row squared norms.
struct Operand {
    __nv_bfloat16 *planes;
    float *sumsq;
    size_t bytes;
};

static Operand carve_operand(void *buf, int64_t rows, int64_t dim, int P) {
    Carver c(buf);
    Operand o;
    o.planes = c.take<__nv_bfloat16>(static_cast<size_t>(P) * rows * gemm::pad_k(dim));
    o.sumsq = c.take<float>(rows);
    o.bytes = c.total();
    return o;
}

static bool split_ok(int split) { return split == AGRL_SPLIT_BF16X2 || split == AGRL_SPLIT_BF16X3; }
static bool metric_ok(int metric) { return metric == AGRL_METRIC_EUCLIDEAN || metric == AGRL_METRIC_COSINE; }

}  // namespace agrl

using namespace agrl;

extern "C" size_t agrl_distance_operand_bytes(int64_t rows, int64_t dim, int split) {
    if (rows < 0 || dim < 1 || !split_ok(split)) return 0;
    return carve_operand(nullptr, rows, dim, split).bytes;
}

extern "C" int agrl_distance_prepare_operand_dev(const float *x, int64_t ld, int64_t rows, int64_t dim,
                                                 int metric, int split, void *operand, size_t operand_bytes,
                                                 void *stream) {
    if (!x || rows < 0 || dim < 1 || ld < dim || !metric_ok(metric) || !split_ok(split)) return AGRL_E_INVALID;
    if (rows > (1 << 30) || dim > (1 << 24)) return AGRL_E_UNSUPPORTED;
    int rc = agrl_device_ok();
    if (rc) return rc;
    Operand o = carve_operand(operand, rows, dim, split);
    if (!operand || operand_bytes < o.bytes) return AGRL_E_WORKSPACE;
    const int normalize = (metric == AGRL_METRIC_COSINE);
    gemm::SplitArgs sa{x, ld, o.planes, normalize ? nullptr : o.sumsq, rows, static_cast<int>(dim),
                       static_cast<int>(gemm::pad_k(dim)), split, normalize};
    return gemm::launch_split_planes(sa, static_cast<cudaStream_t>(stream));
}

extern "C" int agrl_distance_prepared_dev(const void *q_operand, int64_t num_q, const void *g_operand, int64_t num_g,
                                          int64_t dim, int metric, int split, float *out, int64_t ld_out, void *stream) {
    if (!q_operand || !g_operand || !out || num_q < 0 || num_g < 0 || dim < 1 || ld_out < num_g) return AGRL_E_INVALID;
    if (!metric_ok(metric) || !split_ok(split)) return AGRL_E_INVALID;
    if (num_q > (1 << 30) || num_g > (1 << 30) || dim > (1 << 24)) return AGRL_E_UNSUPPORTED;
    int rc = agrl_device_ok();
    if (rc) return rc;
    if (num_q == 0 || num_g == 0) return AGRL_OK;
    Operand q = carve_operand(const_cast<void *>(q_operand), num_q, dim, split);
    Operand g = carve_operand(const_cast<void *>(g_operand), num_g, dim, split);
    const int kp = static_cast<int>(gemm::pad_k(dim));
    CUtensorMap map_q, map_g;
    if ((rc = gemm::make_plane_tensor_map(&map_q, q.planes, num_q, kp, split, gemm::BM, num_q))) return rc;
    const bool pair = option(kOptGemmPair) != 0;
    if ((rc = gemm::make_plane_tensor_map(&map_g, g.planes, num_g, kp, split, pair ? 64 : 128, num_g))) return rc;
    gemm::EpiDistance epi{q.sumsq, g.sumsq, out, ld_out, metric};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nq = static_cast<int>(num_q), ng = static_cast<int>(num_g);
    if (split == AGRL_SPLIT_BF16X3)
        return pair ? gemm::launch_pair_gemm<3, 128, true>(map_q, map_g, nq, ng, kp, epi, st)
                    : gemm::launch_split_gemm<3, 128, true>(map_q, map_g, nq, ng, kp, epi, st);
    return pair ? gemm::launch_pair_gemm<2, 128, true>(map_q, map_g, nq, ng, kp, epi, st)
                : gemm::launch_split_gemm<2, 128, true>(map_q, map_g, nq, ng, kp, epi, st);
}

extern "C" size_t agrl_distance_workspace_bytes(int64_t num_q, int64_t num_g, int64_t dim, int split) {
    if (num_q < 0 || num_g < 0 || dim < 1 || !split_ok(split)) return 0;
    return agrl_distance_operand_bytes(num_q, dim, split) + agrl_distance_operand_bytes(num_g, dim, split);
}

extern "C" int agrl_distance_dev(const float *q, int64_t ld_q, const float *g, int64_t ld_g,
                                 float *out, int64_t ld_out, int64_t num_q, int64_t num_g, int64_t dim,
                                 int metric, int split, void *ws, size_t ws_bytes, void *stream) {
    if (!q || !g || !out) return AGRL_E_INVALID;
    if (num_q < 0 || num_g < 0 || dim < 1 || ld_q < dim || ld_g < dim || ld_out < num_g) return AGRL_E_INVALID;
    if (!metric_ok(metric) || !split_ok(split)) return AGRL_E_INVALID;
    int rc = agrl_device_ok();
    if (rc) return rc;
    if (num_q == 0 || num_g == 0) return AGRL_OK;
    const size_t qb = agrl_distance_operand_bytes(num_q, dim, split), gb = agrl_distance_operand_bytes(num_g, dim, split);
    if (!ws || ws_bytes < qb + gb) return AGRL_E_WORKSPACE;
    char *qo = static_cast<char *>(ws), *go = qo + qb;
    if ((rc = agrl_distance_prepare_operand_dev(q, ld_q, num_q, dim, metric, split, qo, qb, stream))) return rc;
    if ((rc = agrl_distance_prepare_operand_dev(g, ld_g, num_g, dim, metric, split, go, gb, stream))) return rc;
    return agrl_distance_prepared_dev(qo, num_q, go, num_g, dim, metric, split, out, ld_out, stream);
}
