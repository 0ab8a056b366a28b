// host_numa.h -- which CPUs sit next to a GPU (shared by hostbuf.cu and restore_pipe.cu; not part of the ABI)
#pragma once
#include <cuda_runtime.h>
#include <ctype.h>
#include <sched.h>
#include <stdio.h>
#include <string.h>

namespace nvrx {

// CPUs of the NUMA node the GPU hangs off (empty set if sysfs does not say): the slot is first-touched from
// those CPUs so its pages land in the DRAM next to the GPU's PCIe root port and the drain does not cross sockets.
inline bool numa_cpus_of_device(int device, cpu_set_t* set) {
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof(bdf), device) != cudaSuccess) return false;
    for (char* c = bdf; *c; ++c) *c = static_cast<char>(tolower(*c));
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bdf);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    int node = -1;
    const int got = fscanf(f, "%d", &node);
    fclose(f);
    if (got != 1 || node < 0) return false;
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return false;
    char list[4096] = {0};
    const bool ok = fgets(list, sizeof(list), f) != nullptr;
    fclose(f);
    if (!ok) return false;
    CPU_ZERO(set);
    int count = 0;
    for (char* tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int lo = 0, hi = 0;
        const int n = sscanf(tok, "%d-%d", &lo, &hi);
        if (n == 1) hi = lo;
        if (n < 1) continue;
        for (int c = lo; c <= hi && c < CPU_SETSIZE; ++c) {
            CPU_SET(c, set);
            ++count;
        }
    }
    return count > 0;
}


}  // namespace nvrx
