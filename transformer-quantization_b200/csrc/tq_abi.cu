// Library-info entry points of the C ABI (include/tq_b200.h).
#include "tq_common.cuh"

extern "C" {

int tq_version(void) { return 1; }

const char* tq_error_string(int code) {
    switch (code) {
        case TQ_OK: return "ok";
        case TQ_EINVAL: return "tq: invalid argument";
        case TQ_EALIGN: return "tq: pointer alignment requirement not met";
        case TQ_EWORKSPACE: return "tq: workspace too small";
        case TQ_EUNSUPPORTED: return "tq: unsupported configuration";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "tq: unknown error";
}

int tq_device_sm_count(void) { return tq::sm_count(); }

}  // extern "C"
