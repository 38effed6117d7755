/* A C host of the C-ABI (include/la3d.h): what a non-Python pipeline stage links against.
 *
 *   gcc -std=c99 -Iinclude examples/c_host.c -Llabelany3d_b200/lib -lla3d_sm100a -Wl,-rpath,$PWD/labelany3d_b200/lib -o c_host
 *
 * Without arguments it only exercises the calls that need no GPU (version, sizes, argument checks); the
 * device calls take plain device pointers from cudaMalloc / the host's own allocator and a cudaStream_t
 * passed as la3d_stream_t - see INTEGRATION.md section 4 for the whole step. */
#include <stdio.h>
#include <string.h>

#include "la3d.h"

int main(void) {
  const int B = 256, I = 8, H = 480, W = 640;
  printf("la3d version %d\n", la3d_version());
  if (la3d_version() != LA3D_VERSION) {
    fprintf(stderr, "header %d and library %d differ\n", LA3D_VERSION, la3d_version());
    return 1;
  }
  printf("workspace for %d x %d masks of %d x %d: %zu bytes; %zu chunks per plane; preparation %zu bytes\n", B, I, H, W,
         la3d_fit_workspace_bytes(B, I, H, W), la3d_chunks_per_plane(H, W), la3d_prep_bytes(B, I));
  /* argument errors come back as codes with a message, before anything touches the device */
  int rc = la3d_fit_boxes(NULL, NULL, NULL, NULL, B, I, H, W, 1, LA3D_METHOD_PCA, 0, 0u, 0u, NULL, 0, NULL, 0, NULL);
  printf("null pointers -> %d (%s)\n", rc, la3d_last_error());
  if (rc != LA3D_EINVAL || strlen(la3d_last_error()) == 0) return 2;
  la3d_sink sink;
  memset(&sink, 0, sizeof sink);
  printf("record = %d scalars, sink describes up to %d destinations\n", LA3D_REC, (int)(sizeof sink.records / sizeof sink.records[0]));
  return 0;
}
