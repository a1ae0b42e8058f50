/* cabi_smoke.c — TEST INFRASTRUCTURE: include/rxb.h must be plain C (C99, -pedantic) and librxb.so must link and answer
 * from a C program with no C++ or torch in sight: the drop-in boundary is a C ABI. */
#include <stdio.h>
#include "rxb.h"
int main(void) {
  printf("version %d\n", rxb_version());
  rxb_conv_desc d; (void)d;
  rxb_dn121_config c = {8, 256, 256, 1108, 1e-5f, 0.1f};
  printf("params %lld ws %zu jpeg_ws %zu\n", (long long)rxb_dn121_param_count(&c), rxb_dn121_workspace_bytes(&c, 1), rxb_jpeg_decode_workspace_bytes(768, 512, 512));
  int rc = rxb_stats_accumulate(0, 0, 1, 512, 512, 6, 0, 1, 0, 0, 0, 0);
  printf("rc %d: %s\n", rc, rxb_last_error());
  return 0;
}
