/*
 * Per-config compile-time constants for building the reference's binary decoder FOR THE GPU
 * (baseline/build_ref_gpu.sh passes -DREF_*; force-included with -include so that the reference's own define.cuh
 * is skipped by its include guard).  Values per SURVEY Appendix A — the committed B/define.cuh is internally
 * inconsistent.  The CPU-shim twin is oracle/shim/define_override.cuh; it cannot be used here because
 * its directory holds the fake cuda_runtime.h of the shim.  BASELINE INFRASTRUCTURE ONLY.
 */
#ifndef _DEFINE_H_
#define _DEFINE_H_
#include <stdlib.h>
#include <stdio.h>
#include <math.h>
#include <string.h>
#include <memory.h>
#include <time.h>
#include "struct.cuh"
#include "Simulation.cuh"
#include <cuda_runtime.h>
#include <device_launch_parameters.h>

#define J REF_J
#define L REF_L
#define Z REF_Z
#define CW_Len (REF_L * REF_Z)
#define msgLen ((REF_L - REF_J) * REF_Z)
#define parLen (REF_J * REF_Z)
#define PN_Message 0
#define minWeight_checknode 1
#define maxWeight_checknode 25
#define minWeight_variablenode 1
#define maxWeight_variablenode 15
#define rate ((float)msgLen / CW_Len)
#define maxIT REF_MAXIT
#define decoder_method 0
#define ix_define 173
#define iy_define 173
#define iz_define 173
#define Add_noise 1
#define snrtype REF_SNRTYPE
#define startSNR 0
#define stepSNR 1.0
#define stopSNR 0
#define leastErrorFrames 50
#define leastTestFrames REF_LEAST_TEST_FRAMES
#define displayStep 1000000000
#define MaxThreadPerBlock 1024
#define PI (3.1415926)
#define CPU_GPU 1
#define Num_Frames_OneTime REF_F
#define Message_CW 0
#endif
