// kbs_net_tc.cu -- tcgen05 3xTF32 datapath (placeholder until the kernel lands; fails loudly, never falls back).
#include "kbs_common.cuh"

int kbs_tc_pack(kbs_handle*, int, cudaStream_t) { return KBS_E_STATE; }
int kbs_tc_trunk(kbs_handle*, int, const float*, int64_t, float*, const uint8_t*, float*, int64_t, cudaStream_t) {
  return KBS_E_STATE;
}
