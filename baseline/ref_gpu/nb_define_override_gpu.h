/*
 * Per-config compile-time constants for building the reference's NON-BINARY simulator (myNBLDPC/src, main.cu
 * included, all sources unchanged) FOR THE GPU: baseline/build_ref_nb_gpu.sh passes -DNBREF_*; force-included with
 * -include so that the reference's own include/define.h is skipped by its include guard.  CPU_GPU = 1 selects its
 * frame-at-a-time GPU decoders (Decoding_EMS_GPU / Decoding_TMM_GPU, src/Decode_GPU.cu:138,706).
 * BASELINE INFRASTRUCTURE ONLY.
 */
#ifndef _DEFINE_H_
#define _DEFINE_H_
#include <stdlib.h>
#include <stdio.h>
#include <math.h>
#include <string.h>
#include <memory.h>
#include <time.h>
#include "struct.h"
#include "Simulation.h"
#include <cuda_runtime.h>
#include <device_launch_parameters.h>

#define Matrixfile NBREF_MATRIX
#define Constellationfile NBREF_CONST
#define n_QAM NBREF_NQAM
#define GFQ NBREF_GFQ
#define maxdc NBREF_MAXDC
#define maxdv NBREF_MAXDV
#define THREAD_NUM 1
#define EMS_NM NBREF_NM
#define EMS_NC NBREF_NC
#define maxIT NBREF_MAXIT
#define decoder_method NBREF_METHOD
#define ix_define 173
#define iy_define 173
#define iz_define 173
#define Add_noise 1
#define snrtype 0
#define startSNR NBREF_SNR
#define stepSNR 1.0
#define stopSNR NBREF_SNR
#define leastErrorFrames NBREF_LEAST_ERRORS
#define leastTestFrames NBREF_LEAST_FRAMES
#define displayStep 100000
#define PI (3.1415926)
#define CPU_GPU 1
#endif
