// Development harness: csrc/png.cu under the lock-step emulator (see cuda_emu.h).  Never shipped.
#define SAGE_CUDA_EMU 1
#include "../../sage-3d_official_b200/csrc/api_common.cuh"
std::string& sage_err_slot()
{
    static thread_local std::string err;
    return err;
}
extern "C" const char* sage_last_error(void) { return sage_err_slot().c_str(); }
#include "../../sage-3d_official_b200/csrc/png.cu"
