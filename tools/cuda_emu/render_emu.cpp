// Development harness: the render path (csrc/project.cu, sort.cu, bin.cu, composite.cu, frame.cu and
// the C-ABI layer scene_api.cu) under the lock-step emulator (cuda_emu.h).  Every translation unit is
// compiled separately with g++ -x c++ -DSAGE_CUDA_EMU and linked with this file, which supplies the
// pieces of api.cu the emulator cannot host (device discovery, IPC).  CUDA graphs are stubbed out, so
// sage_render takes its direct-launch path.  Never part of libsage_splat.so.
#define SAGE_CUDA_EMU 1
#include "../../sage-3d_official_b200/csrc/api_common.cuh"

std::string& sage_err_slot()
{
    static thread_local std::string err;
    return err;
}
extern "C" int sage_device_count(void) { return 1; }
extern "C" const char* sage_last_error(void) { return sage_err_slot().c_str(); }
