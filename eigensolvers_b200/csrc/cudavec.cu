// cudavec.cu — single translation unit of libcudavec.so (unity build: the kernels are templates
// shared by the API layer, the solvers and the communication layer).
#include "api.cu"
#include "comm.cu"
#include "peer.cu"
#include "solvers.cu"
