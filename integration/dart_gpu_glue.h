/* dart_gpu_glue.h — the three entry points integration/dart_gpu.patch adds to the reference's main.cpp / Mapping.cpp.
 * Everything else of the binding lives in dart_gpu_glue.cpp, compiled next to the reference's own sources. */
#ifndef DART_GPU_GLUE_H
#define DART_GPU_GLUE_H
bool  DartGpuInit();                 /* after bwa_idx_load + RestoreReferenceInfo (main.cpp:220-222): index -> every GPU's HBM */
bool  DartGpuActive();
void *DartGpuReadMapping(void *arg); /* the body of ReadMapping() (Mapping.cpp:580-680) with the per-read loop on the GPU   */
void  DartGpuShutdown();
#endif
