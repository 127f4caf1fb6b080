// gds_probe.cu -- can this box write a device buffer to a file with GPUDirect Storage (cuFile), and how fast is it next
// to the pinned-staging path the exporter uses (cudaMemcpyAsync D2H + pwrite)?   SURVEY.md 8(f) rank 1.
//   nvcc -O2 -o gds_probe gds_probe.cu -lcufile ; ./gds_probe <dir> [GiB]
#include <cuda_runtime.h>
#include <cufile.h>
#include <fcntl.h>
#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
    const std::string dir = argc > 1 ? argv[1] : ".";
    const size_t gib = argc > 2 ? (size_t)atoi(argv[2]) : 2, bytes = gib << 30, hop = 32u << 20;
    void *d = nullptr; if (cudaMalloc(&d, bytes) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
    cudaMemset(d, 0x5a, bytes);
    // ---- baseline: pinned staging, double buffered
    {
        const std::string path = dir + "/gds_probe_plain.bin";
        int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        void *hb[2]; cudaMallocHost(&hb[0], hop); cudaMallocHost(&hb[1], hop);
        cudaStream_t s; cudaStreamCreate(&s); cudaEvent_t ev[2]; cudaEventCreate(&ev[0]); cudaEventCreate(&ev[1]);
        double t0 = now();
        const size_t n = bytes / hop;
        cudaMemcpyAsync(hb[0], d, hop, cudaMemcpyDeviceToHost, s); cudaEventRecord(ev[0], s);
        bool ok = fd >= 0;
        for (size_t c = 0; c < n && ok; c++) {
            cudaEventSynchronize(ev[c & 1]);
            if (c + 1 < n) { cudaMemcpyAsync(hb[(c + 1) & 1], (char *)d + (c + 1) * hop, hop, cudaMemcpyDeviceToHost, s); cudaEventRecord(ev[(c + 1) & 1], s); }
            ok = pwrite(fd, hb[c & 1], hop, (off_t)(c * hop)) == (ssize_t)hop;
        }
        double t1 = now(); if (fd >= 0) { fsync(fd); close(fd); } double t2 = now();
        printf("pinned staging + pwrite: %s  %.2f GB/s into the page cache, %.2f GB/s incl. fsync\n", ok ? "ok" : "FAILED", bytes / (t1 - t0) / 1e9, bytes / (t2 - t0) / 1e9);
        unlink(path.c_str());
        // D2H only (no file): the PCIe ceiling of this box
        t0 = now();
        for (size_t c = 0; c < n; c++) cudaMemcpyAsync(hb[c & 1], (char *)d + c * hop, hop, cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        printf("D2H only, one stream: %.2f GB/s\n", bytes / (now() - t0) / 1e9);
        cudaStream_t s2; cudaStreamCreate(&s2); void *hb2[2]; cudaMallocHost(&hb2[0], hop); cudaMallocHost(&hb2[1], hop);
        t0 = now();
        for (size_t c = 0; c < n; c += 2) { cudaMemcpyAsync(hb[(c >> 1) & 1], (char *)d + c * hop, hop, cudaMemcpyDeviceToHost, s); if (c + 1 < n) cudaMemcpyAsync(hb2[(c >> 1) & 1], (char *)d + (c + 1) * hop, hop, cudaMemcpyDeviceToHost, s2); }
        cudaStreamSynchronize(s); cudaStreamSynchronize(s2);
        printf("D2H only, two streams: %.2f GB/s\n", bytes / (now() - t0) / 1e9);
    }
    // ---- cuFile
    CUfileError_t st = cuFileDriverOpen();
    if (st.err != CU_FILE_SUCCESS) { printf("cuFileDriverOpen failed: err %d (cuda %d) -- no GDS on this box\n", (int)st.err, (int)st.cu_err); return 0; }
    {
        const std::string path = dir + "/gds_probe_cufile.bin";
        int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_DIRECT, 0644);
        if (fd < 0) { printf("open(O_DIRECT) failed on %s: %s\n", dir.c_str(), strerror(errno)); cuFileDriverClose(); return 0; }
        CUfileDescr_t desc; memset(&desc, 0, sizeof desc); desc.handle.fd = fd; desc.type = CU_FILE_HANDLE_TYPE_OPAQUE_FD;
        CUfileHandle_t fh;
        st = cuFileHandleRegister(&fh, &desc);
        if (st.err != CU_FILE_SUCCESS) { printf("cuFileHandleRegister failed: err %d -- file system not GDS-capable (compat mode off?)\n", (int)st.err); close(fd); unlink(path.c_str()); cuFileDriverClose(); return 0; }
        st = cuFileBufRegister(d, bytes, 0);
        printf("cuFileBufRegister: %s\n", st.err == CU_FILE_SUCCESS ? "ok" : "failed (unregistered buffers still work through bounce buffers)");
        double t0 = now(); ssize_t tot = 0;
        for (size_t off = 0; off < bytes; off += hop) { ssize_t w = cuFileWrite(fh, d, hop, (off_t)off, (off_t)off); if (w < 0) { printf("cuFileWrite failed at %zu: %zd\n", off, w); break; } tot += w; }
        double t1 = now(); fsync(fd); double t2 = now();
        printf("cuFileWrite: %zd bytes  %.2f GB/s, %.2f GB/s incl. fsync\n", tot, tot / (t1 - t0) / 1e9, tot / (t2 - t0) / 1e9);
        cuFileBufDeregister(d); cuFileHandleDeregister(fh); close(fd); unlink(path.c_str());
    }
    cuFileDriverClose();
    return 0;
}
