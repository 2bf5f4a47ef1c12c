// tvc_rollout.cu -- fused actor-MLP rollout (config 4).  Placeholder until the tcgen05 kernel lands.
#include "tvc_internal.h"

void tvc_rollout_free(tvc_handle *h) { (void)h; }

extern "C" int tvc_rollout(tvc_handle *h, const tvc_actor_weights *w, int32_t T, const tvc_rollout_io *io, tvc_stream stream) {
    (void)h; (void)w; (void)T; (void)io; (void)stream;
    tvc_set_err("tvc_rollout: not implemented in this build");
    return TVC_E_STATE;
}
