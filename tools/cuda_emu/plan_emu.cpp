// Development harness: csrc/plan.cu under the lock-step emulator (cuda_emu.h).
#define SAGE_CUDA_EMU 1
#include "../../sage-3d_official_b200/csrc/plan.cu"

std::string& sage_err_slot()
{
    static thread_local std::string err;
    return err;
}
extern "C" int sage_device_count(void) { return 1; }
extern "C" const char* sage_last_error(void) { return sage_err_slot().c_str(); }
