// libb200rl: version / error reporting / workspace layout.  ABI: include/b200rl.h
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_last_error[1024] = "";
long long g_b200rl_launches = 0;

void b200rl_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

int b200rl_validate_net(const b200rl_net* net, const char* name, bool is_actor) {
    B200RL_REQUIRE(net != nullptr, "%s: net is NULL", name);
    B200RL_REQUIRE(net->num_linear >= 1 && net->num_linear <= B200RL_MAX_LINEAR,
                   "%s: num_linear=%d outside [1, %d]", name, net->num_linear, B200RL_MAX_LINEAR);
    B200RL_REQUIRE(net->activation == B200RL_ACT_GELU || net->activation == B200RL_ACT_RELU,
                   "%s: unknown activation %d", name, net->activation);
    for (int l = 0; l <= net->num_linear; ++l)
        B200RL_REQUIRE(net->dims[l] >= 1 && net->dims[l] <= 1024, "%s: dims[%d]=%d outside [1, 1024]", name, l,
                       net->dims[l]);
    for (int l = 0; l < net->num_linear; ++l)
        B200RL_REQUIRE(net->weight[l] && net->bias[l], "%s: weight/bias pointer of layer %d is NULL", name, l);
    B200RL_REQUIRE((net->state_avg == nullptr) == (net->state_std == nullptr),
                   "%s: state_avg and state_std must both be given or both NULL", name);
    if (is_actor) B200RL_REQUIRE(net->action_std_log != nullptr, "%s: actor needs action_std_log", name);
    return 0;
}

extern "C" {

const char* b200rl_version(void) { return "b200rl 0.1.0 (sm_100a)"; }
const char* b200rl_last_error(void) { return g_last_error; }
int64_t b200rl_launch_count(void) { return (int64_t)g_b200rl_launches; }

int64_t b200rl_grad_numel(const b200rl_net* actor, const b200rl_net* critic) {
    if (!actor || !critic) return -1;
    return ((b200rl_net_numel(actor) + 3) & ~(int64_t)3) + b200rl_net_numel(critic);  // critic segment 16-byte aligned
}
int64_t b200rl_workspace_grad_offset(void) { return B200RL_WS_HEADER_BYTES; }
int64_t b200rl_workspace_bytes(const b200rl_net* actor, const b200rl_net* critic) {
    int64_t n = b200rl_grad_numel(actor, critic);
    if (n < 0) return -1;
    const int64_t stride = (n + 63) & ~(int64_t)63;  // two gradient buffers (the persistent update kernel alternates)
    return B200RL_WS_HEADER_BYTES + 2 * stride * 4;
}

}  // extern "C"
