// dps_plan.h -- the host-side plan object behind the opaque ``dps_plan*`` of include/dps.h.
#pragma once
#include <vector>
#include "dps_internal.h"

struct dps_plan {
    dps_topology topo;                    // scalar members valid; pointers re-targeted to the vectors below
    std::vector<int32_t> ctrl;
    std::vector<uint8_t> step_cfg;
    std::vector<int32_t> slot_vary;
    std::vector<dps_channel> rec;
    std::vector<int> seg_step, seg_cfg;   // configuration change points
    int g_template;
    bool any_ctrl, any_vary;
    int cmask;                            // common controller flag set of all generators (0 or 7), or -1 if mixed
    int pmode;                            // 1 = machine constants in registers, 2 = generic slot table
    bool wide;                            // more than MAX_G generators: dps_step_wide_kernel
    dps::WideParams wbase;
    dps::FwdParams base;                       // everything that does not depend on the call
    mutable long long last_P = 0;              // batch size of the latest forward call (dps_plan_info reports its kernel)
};
