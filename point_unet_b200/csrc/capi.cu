// capi.cu -- library-wide C-ABI entry points and globals of libpointunet_b200.so
#include "common.cuh"

namespace pu {
thread_local int g_last_cuda_error = 0;
unsigned long long g_launch_count = 0;
}  // namespace pu

extern "C" {

const char *pu_version(void) { return "pointunet_b200 0.1 (sm_100a)"; }

int pu_last_cuda_error(void) { return pu::g_last_cuda_error; }

unsigned long long pu_launch_count(void) { return __atomic_load_n(&pu::g_launch_count, __ATOMIC_RELAXED); }

// CRC32C (Castagnoli, reflected 0x82F63B78), slicing-by-8 on the host: the checksum TensorFlow's checkpoint format stores
// per tensor and per index block (point_unet_b200/tf_checkpoint.py; pure host-side I/O helper, no device work).
unsigned int pu_crc32c(const void *data, size_t n, unsigned int crc) {
    static unsigned int T[8][256];
    static bool ready = false;
    if (!ready) {
        for (int i = 0; i < 256; ++i) {
            unsigned int c = (unsigned int)i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
            T[0][i] = c;
        }
        for (int i = 0; i < 256; ++i)
            for (int t = 1; t < 8; ++t) T[t][i] = (T[t - 1][i] >> 8) ^ T[0][T[t - 1][i] & 0xFF];
        __atomic_store_n(&ready, true, __ATOMIC_RELEASE);
    }
    const unsigned char *p = (const unsigned char *)data;
    unsigned int c = crc ^ 0xFFFFFFFFu;
    while (n >= 8) {
        const unsigned int lo = ((unsigned int)p[0] | ((unsigned int)p[1] << 8) | ((unsigned int)p[2] << 16) | ((unsigned int)p[3] << 24)) ^ c;
        c = T[7][lo & 0xFF] ^ T[6][(lo >> 8) & 0xFF] ^ T[5][(lo >> 16) & 0xFF] ^ T[4][lo >> 24] ^ T[3][p[4]] ^ T[2][p[5]] ^
            T[1][p[6]] ^ T[0][p[7]];
        p += 8;
        n -= 8;
    }
    while (n--) c = T[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

}  // extern "C"
