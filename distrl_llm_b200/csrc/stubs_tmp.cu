// TEMPORARY bring-up stubs (removed once attention.cu / model.cu land)
#include "common.cuh"
using namespace b200rl;
#define STUB(name, ...) extern "C" int name(__VA_ARGS__) { return set_error(B200RL_ERR_UNSUPPORTED, #name " not built yet"); }
extern "C" long long b200rl_model_workspace_bytes(const void*) { return 0; }
extern "C" long long b200rl_model_lora_numel(const void*) { return 0; }
STUB(b200rl_model_create, const void*, const void*, const void*, const void*, const void*, const void*, float*, float*, void*, long long, void**)
STUB(b200rl_model_destroy, void*)
STUB(b200rl_model_sync_lora, void*, void*)
STUB(b200rl_model_microbatch, void*, const int*, const int*, const int*, const double*, float*, double*, int, int, int, int, int, int, void*)
