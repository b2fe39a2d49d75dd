/* tests/c/abi_check.c -- the C ABI from plain C: include/marlnav_b200.h must compile as C99, every
 * declared entry point must link, struct sizes must agree with the library's, and argument errors
 * must come back as codes with a message (no GPU needed). */
#include <stdio.h>
#include <string.h>
#include "../../include/marlnav_b200.h"

int main(void) {
    void* entry[] = {(void*)marlnav_abi_version, (void*)marlnav_last_error, (void*)marlnav_obs_size,
                     (void*)marlnav_device_count, (void*)marlnav_counter_add, (void*)marlnav_init_f32,
                     (void*)marlnav_observe_f32, (void*)marlnav_step_f32, (void*)marlnav_step_call_f32,
                     (void*)marlnav_host_pipe_create, (void*)marlnav_host_pipe_destroy, (void*)marlnav_step_host_f32,
                     (void*)marlnav_step_launch_info, (void*)marlnav_actor_sample_f32, (void*)marlnav_act_step_f32,
                     (void*)marlnav_critic_value_f32, (void*)marlnav_discounted_returns_f64,
                     (void*)marlnav_rollout_last_error};
    size_t i;
    for (i = 0; i < sizeof entry / sizeof entry[0]; ++i) if (!entry[i]) return 10;
    if (marlnav_abi_version() != MARLNAV_ABI_VERSION) return 1;
    if (marlnav_sizeof_env_params() != sizeof(marlnav_env_params)) return 2;
    if (marlnav_sizeof_reset_spec() != sizeof(marlnav_reset_spec)) return 3;
    if (marlnav_sizeof_io_transform() != sizeof(marlnav_io_transform)) return 4;
    if (marlnav_sizeof_actor_spec() != sizeof(marlnav_actor_spec)) return 5;
    if (marlnav_sizeof_step_call() != sizeof(marlnav_step_call)) return 6;
    if (marlnav_obs_size(3, 3) != 12 || marlnav_obs_size(8, 16) != 48) return 7;
    {
        marlnav_env_params p;
        marlnav_reset_spec rs;
        marlnav_step_call call;
        float dummy[4];
        memset(&p, 0, sizeof p); memset(&rs, 0, sizeof rs); memset(&call, 0, sizeof call);
        p.struct_size = sizeof p; p.num_envs = 4; p.num_agents = 3; p.num_obstacles = 3;
        rs.struct_size = (uint32_t)sizeof rs - 8;          /* a binding built against an older layout */
        rs.alias_first_step = 1;
        call.struct_size = sizeof call; call.params = &p; call.reset = &rs;
        call.states = call.obstacles = call.target = call.step_num = dummy; call.terminates = (uint8_t*)dummy;
        call.actions = dummy; call.obs = call.rewards = dummy; call.terminated = call.truncated = (uint8_t*)dummy;
        call.stats = (unsigned long long*)dummy;
        if (marlnav_step_call_f32(&call) != MARLNAV_ERR_BAD_ARG) return 8;
        if (!strstr(marlnav_last_error(), "marlnav_reset_spec.struct_size")) return 9;
    }
    printf("abi %d ok: env_params %zu, reset_spec %zu, io_transform %zu, actor_spec %zu, step_call %zu bytes\n",
           marlnav_abi_version(), sizeof(marlnav_env_params), sizeof(marlnav_reset_spec), sizeof(marlnav_io_transform),
           sizeof(marlnav_actor_spec), sizeof(marlnav_step_call));
    return 0;
}
