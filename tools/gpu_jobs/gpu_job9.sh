#!/bin/bash
export CCD_MHSA_FWD_VARIANT=2
ROUND=r01b PKERNELS="gemm_umma_persistent_kernel mhsa_bwd_kernel mhsa_fwd_persistent_kernel" PCOUNT=6 bash tools/gpu_profile.sh 2>&1 | tail -12
