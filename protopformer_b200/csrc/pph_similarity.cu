// pph_similarity_fwd: dispatch between the FP32-FMA (CUDA core) mode and the tcgen05 tensor-core modes.
// There is no CPU path: every mode launches sm_100a kernels.
#include "pph_common.cuh"

namespace pph {
int similarity_fwd_simt(int act_fn, float eps, int B, int K, int D, int P, int Pg,
                        const float* Zs, const float* Zc, const float* z2s, const float* z2c,
                        const float* Pl, const float* Pgl, const float* p2l, const float* p2g,
                        float* dmin_l, int32_t* argmin_l, float* act_l, float* dmin_g, float* act_g,
                        float* dist_map, float* act_map, cudaStream_t st);
int similarity_fwd_tc(int mode, int act_fn, float eps, int B, int K, int D, int P, int Pg,
                      const float* z2s, const float* z2c,
                      const uint16_t* Zs_hi, const uint16_t* Zs_lo, const uint16_t* Zc_hi, const uint16_t* Zc_lo,
                      const float* p2l, const float* p2g,
                      const uint16_t* Pl_hi, const uint16_t* Pl_lo, const uint16_t* Pg_hi, const uint16_t* Pg_lo,
                      float* dmin_l, int32_t* argmin_l, float* act_l, float* dmin_g, float* act_g, cudaStream_t st);
int similarity_plan(int mode, int B, int K, int D, int P, int Pg, int sms, int* out, int* coverage);
}  // namespace pph

extern "C" int pph_similarity_plan(int mode, int B, int K, int D, int P, int Pg, int sms, int* out, int* coverage) {
    using namespace pph;
    PPH_REQUIRE(mode == PPH_MODE_BF16X3 || mode == PPH_MODE_BF16, PPH_EINVAL, "pph_similarity_plan: tensor-core modes only");
    PPH_REQUIRE(B >= 1 && P >= 1 && Pg >= 0 && D % 64 == 0 && D >= 64 && D <= 512 && K >= 1 && K <= 256, PPH_EUNSUP,
                "pph_similarity_plan: shape outside the tcgen05 build range (B=%d K=%d D=%d P=%d Pg=%d)", B, K, D, P, Pg);
    return similarity_plan(mode, B, K, D, P, Pg, sms, out, coverage);
}

extern "C" int pph_similarity_fwd(int mode, int act_fn, float eps, int B, int K, int D, int P, int Pg,
                                  const float* Zs, const float* Zc, const float* z2s, const float* z2c,
                                  const uint16_t* Zs_hi, const uint16_t* Zs_lo, const uint16_t* Zc_hi,
                                  const uint16_t* Zc_lo,
                                  const float* Pl, const float* Pgl, const float* p2l, const float* p2g,
                                  const uint16_t* Pl_hi, const uint16_t* Pl_lo, const uint16_t* Pg_hi,
                                  const uint16_t* Pg_lo,
                                  float* dmin_l, int32_t* argmin_l, float* act_l, float* dmin_g, float* act_g,
                                  float* dist_map, float* act_map, pph_stream_t stream) {
    using namespace pph;
    PPH_REQUIRE(B >= 0 && K >= 1 && D >= 1 && P >= 1 && Pg >= 0, PPH_EINVAL,
                "pph_similarity_fwd: bad dims B=%d K=%d D=%d P=%d Pg=%d", B, K, D, P, Pg);
    PPH_REQUIRE(act_fn == PPH_ACT_LOG || act_fn == PPH_ACT_LINEAR, PPH_EINVAL, "pph_similarity_fwd: bad act_fn %d",
                act_fn);
    if (B == 0) return 0;
    if (mode == PPH_MODE_FP32_FMA) {
        PPH_REQUIRE(Zs && z2s && Pl && p2l && dmin_l && argmin_l && act_l, PPH_EINVAL,
                    "pph_similarity_fwd(fp32): null local pointer");
        PPH_REQUIRE(Pg == 0 || (Zc && z2c && Pgl && p2g && dmin_g && act_g), PPH_EINVAL,
                    "pph_similarity_fwd(fp32): null global pointer");
        return similarity_fwd_simt(act_fn, eps, B, K, D, P, Pg, Zs, Zc, z2s, z2c, Pl, Pgl, p2l, p2g, dmin_l, argmin_l,
                                   act_l, dmin_g, act_g, dist_map, act_map, as_stream(stream));
    }
    PPH_REQUIRE(mode == PPH_MODE_BF16X3 || mode == PPH_MODE_BF16, PPH_EINVAL, "pph_similarity_fwd: bad mode %d", mode);
    PPH_REQUIRE(!dist_map && !act_map, PPH_EUNSUP,
                "pph_similarity_fwd: the tensor-core modes never materialise the (B,P,K) map; use PPH_MODE_FP32_FMA");
    return similarity_fwd_tc(mode, act_fn, eps, B, K, D, P, Pg, z2s, z2c, Zs_hi, Zs_lo, Zc_hi, Zc_lo, p2l, p2g, Pl_hi,
                             Pl_lo, Pg_hi, Pg_lo, dmin_l, argmin_l, act_l, dmin_g, act_g, as_stream(stream));
}
