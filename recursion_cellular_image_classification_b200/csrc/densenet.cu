// densenet.cu — DenseNet-121 executor (placeholder entry points while the executor is brought up).
#include "common.cuh"
extern "C" {
int64_t rxb_dn121_param_count(const rxb_dn121_config*) { return 0; }
int64_t rxb_dn121_buffer_count(const rxb_dn121_config*) { return 0; }
size_t rxb_dn121_workspace_bytes(const rxb_dn121_config*, int) { return 0; }
int rxb_dn121_create(const rxb_dn121_config*, float*, float*, float*, float*, void*, size_t, int, rxb_dn121**) {
  return rxb::set_error(RXB_ERR_UNSUPPORTED, "dn121 executor not built yet");
}
void rxb_dn121_destroy(rxb_dn121*) {}
int rxb_dn121_sync_weights(rxb_dn121*, rxb_stream_t) { return rxb::set_error(RXB_ERR_UNSUPPORTED, "nyi"); }
int rxb_dn121_forward(rxb_dn121*, const void*, float*, int, rxb_stream_t) { return rxb::set_error(RXB_ERR_UNSUPPORTED, "nyi"); }
int rxb_dn121_num_phases(void) { return 1; }
int rxb_dn121_phase_grad_range(const rxb_dn121*, int, int64_t*, int64_t*) { return rxb::set_error(RXB_ERR_UNSUPPORTED, "nyi"); }
int rxb_dn121_train_step(rxb_dn121*, const void*, const int64_t*, int, float*, int, rxb_stream_t) { return rxb::set_error(RXB_ERR_UNSUPPORTED, "nyi"); }
int rxb_dn121_sgd(rxb_dn121*, float, float, float, int, float, rxb_stream_t) { return rxb::set_error(RXB_ERR_UNSUPPORTED, "nyi"); }
}
