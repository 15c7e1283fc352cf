// Development harness: the JPEG kernels of csrc/jpeg.cu under the lock-step emulator (cuda_emu.h).
// Build: g++ -std=c++20 -O1 -DSAGE_CUDA_EMU -shared -fPIC -pthread -o /tmp/libjpeg_emu.so jpeg_emu.cpp
#define SAGE_CUDA_EMU 1
#include "../../sage-3d_official_b200/csrc/jpeg.cu"

std::string& sage_err_slot()
{
    static thread_local std::string err;
    return err;
}
extern "C" int sage_device_count(void) { return 1; }
extern "C" const char* sage_last_error(void) { return sage_err_slot().c_str(); }
