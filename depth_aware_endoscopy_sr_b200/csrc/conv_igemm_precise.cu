// The fp32-split ("precise", dasr_set_planes(3)) instantiations of the implicit-GEMM convolution: the same kernel
// template as conv_igemm.cu with PREC = true -- the K loop runs the cross terms of the operand planes, the epilogues
// read plane sums and write plane splits (dasr_internal.h).  A separate translation unit so that the product kernels
// of conv_igemm.cu are compiled exactly as before and both files build in parallel.  Test infrastructure: the
// gradient / forward parity tests run the whole network through these kernels to compare with the fp64 goldens.
#define DASR_CONV_PRECISE_TU
#include "conv_igemm.cu"
