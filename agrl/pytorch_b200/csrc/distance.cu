// distance.cu -- query x gallery distance matrix (torchreid/metrics/distance.py:11-89) on tcgen05.
//
//   euclidean : out[i,j] = (|q_i|^2 + |g_j|^2) - 2 q_i.g_j      (no clamp, no sqrt; distance.py:70-72)
//   cosine    : out[i,j] = 1 - q^_i . g^_j,  x^ = x / max(|x|_2, 1e-12)   (distance.py:86-88)
//
// Pipeline: split.cu turns both operands into bf16 planes (and norms / normalised rows) in one
// pass each, then one persistent tcgen05 GEMM (gemm_sm100.cuh) with the distance epilogue fused.
#include "gemm_sm100.cuh"

namespace agrl {

struct DistWorkspace {
    __nv_bfloat16 *q_planes, *g_planes;
    float *qn, *gn;
    size_t bytes;
};

static DistWorkspace carve_dist(void *ws, int64_t nq, int64_t ng, int64_t dim, int P) {
    Carver c(ws);
    DistWorkspace w;
    const int64_t kp = gemm::pad_k(dim);
    w.q_planes = c.take<__nv_bfloat16>(static_cast<size_t>(P) * nq * kp);
    w.g_planes = c.take<__nv_bfloat16>(static_cast<size_t>(P) * ng * kp);
    w.qn = c.take<float>(nq);
    w.gn = c.take<float>(ng);
    w.bytes = c.total();
    return w;
}

}  // namespace agrl

using namespace agrl;

extern "C" size_t agrl_distance_workspace_bytes(int64_t num_q, int64_t num_g, int64_t dim, int split) {
    if (num_q < 0 || num_g < 0 || dim < 1 || (split != AGRL_SPLIT_BF16X2 && split != AGRL_SPLIT_BF16X3)) return 0;
    return carve_dist(nullptr, num_q, num_g, dim, split).bytes;
}

extern "C" int agrl_distance_dev(const float *q, int64_t ld_q, const float *g, int64_t ld_g,
                                 float *out, int64_t ld_out, int64_t num_q, int64_t num_g, int64_t dim,
                                 int metric, int split, void *ws, size_t ws_bytes, void *stream) {
    if (!q || !g || !out) return AGRL_E_INVALID;
    if (num_q < 0 || num_g < 0 || dim < 1 || ld_q < dim || ld_g < dim || ld_out < num_g) return AGRL_E_INVALID;
    if (metric != AGRL_METRIC_EUCLIDEAN && metric != AGRL_METRIC_COSINE) return AGRL_E_INVALID;
    if (split != AGRL_SPLIT_BF16X2 && split != AGRL_SPLIT_BF16X3) return AGRL_E_INVALID;
    if (num_q > (1 << 30) || num_g > (1 << 30) || dim > (1 << 24)) return AGRL_E_UNSUPPORTED;
    int rc = agrl_device_ok();
    if (rc) return rc;
    if (num_q == 0 || num_g == 0) return AGRL_OK;
    DistWorkspace w = carve_dist(ws, num_q, num_g, dim, split);
    if (!ws || ws_bytes < w.bytes) return AGRL_E_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int kp = static_cast<int>(gemm::pad_k(dim));
    const int normalize = (metric == AGRL_METRIC_COSINE);

    gemm::SplitArgs sq{q, ld_q, w.q_planes, normalize ? nullptr : w.qn, num_q, static_cast<int>(dim), kp, split, normalize};
    gemm::SplitArgs sg{g, ld_g, w.g_planes, normalize ? nullptr : w.gn, num_g, static_cast<int>(dim), kp, split, normalize};
    if ((rc = gemm::launch_split_planes(sq, st))) return rc;
    if ((rc = gemm::launch_split_planes(sg, st))) return rc;

    CUtensorMap map_q, map_g;
    if ((rc = gemm::make_plane_tensor_map(&map_q, w.q_planes, num_q, kp, split, gemm::BM))) return rc;
    if ((rc = gemm::make_plane_tensor_map(&map_g, w.g_planes, num_g, kp, split, 128))) return rc;

    gemm::EpiDistance epi{w.qn, w.gn, out, ld_out, metric};
    if (split == AGRL_SPLIT_BF16X3)
        return gemm::launch_split_gemm<3, 128, true>(map_q, map_g, static_cast<int>(num_q), static_cast<int>(num_g), kp, epi, st);
    return gemm::launch_split_gemm<2, 128, true>(map_q, map_g, static_cast<int>(num_q), static_cast<int>(num_g), kp, epi, st);
}
