/* A host program in plain C on the library's C ABI (include/rtmpc.h): reads the problem description written by
 * examples/dump_qp_desc.py, creates the device-resident QP, solves a batch of (x_init, ref) instances from HOST
 * buffers and prints the packets' first control moves.  No Python, no torch.
 *
 *   python examples/dump_qp_desc.py /tmp/qp_cp.bin
 *   gcc -O2 -std=c99 -Iinclude examples/c_abi_demo.c -o /tmp/c_abi_demo \
 *       -L robust-tracking-mpc-over-lossy-networks_b200/rtmpc_b200 -lrtmpc_b200 \
 *       -Wl,-rpath,$PWD/robust-tracking-mpc-over-lossy-networks_b200/rtmpc_b200
 *   /tmp/c_abi_demo /tmp/qp_cp.bin 4096
 *
 * This is what a maintainer of a compiled host (C, C++, or any language with a C FFI) binds; the reference itself is
 * Python and binds the same entry points through ctypes. */
#define _POSIX_C_SOURCE 199309L
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "rtmpc.h"

static void* read_array(FILE* f, size_t elem, int64_t* count) {
    void* p = NULL;
    if (fread(count, sizeof(int64_t), 1, f) != 1) { fprintf(stderr, "short file\n"); exit(2); }
    if (*count > 0) {
        p = malloc((size_t)*count * elem);
        if (!p || fread(p, elem, (size_t)*count, f) != (size_t)*count) { fprintf(stderr, "short file\n"); exit(2); }
    }
    return p;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s qp.bin [batch]\n", argv[0]); return 2; }
    const int B = argc > 2 ? atoi(argv[2]) : 1024;
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    char magic[8];
    int32_t ints[12];
    double fl[2];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "RTMPCQP1", 8) != 0) { fprintf(stderr, "bad magic\n"); return 2; }
    if (fread(ints, sizeof(int32_t), 12, f) != 12 || fread(fl, sizeof(double), 2, f) != 2) { fprintf(stderr, "short file\n"); return 2; }
    rtmpc_qp_desc d;
    memset(&d, 0, sizeof d);
    d.nx = ints[0]; d.nu = ints[1]; d.N = ints[2]; d.n = ints[3]; d.npad = ints[4]; d.m = ints[5]; d.mpad = ints[6];
    d.np = ints[7]; d.nz = ints[8]; d.nss = ints[9]; d.max_iter = ints[10]; d.min_rows = ints[11];
    d.s_floor = fl[0]; d.sc_b = fl[1];
    int64_t c;
    d.Hs = read_array(f, 8, &c);   d.Hinv = read_array(f, 8, &c); d.G = read_array(f, 8, &c);   d.Y = read_array(f, 8, &c);
    d.Fx = read_array(f, 8, &c);   d.Fr = read_array(f, 8, &c);   d.lo0 = read_array(f, 8, &c); d.up0 = read_array(f, 8, &c);
    d.Lx = read_array(f, 8, &c);   d.Ux = read_array(f, 8, &c);
    d.has_lo = read_array(f, 1, &c); d.has_up = read_array(f, 1, &c);
    d.parC = read_array(f, 8, &c); d.parh = read_array(f, 8, &c); d.Dscale = read_array(f, 8, &c);
    d.Phi = read_array(f, 8, &c);  d.Psi = read_array(f, 8, &c);  d.Kss = read_array(f, 8, &c);
    d.shift = read_array(f, 4, &c);
    fclose(f);

    if (rtmpc_abi_version() != RTMPC_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }
    if (rtmpc_device_count() < 1) { fprintf(stderr, "no CUDA device: %s\n", rtmpc_last_error()); return 1; }
    rtmpc_qp* qp = NULL;
    if (rtmpc_qp_create(&d, &qp) != 0) { fprintf(stderr, "rtmpc_qp_create: %s\n", rtmpc_last_error()); return 1; }

    /* instances: states on a line from the origin towards the reference, target (0.5, 0, 0, 0) */
    const int nx = d.nx, nu = d.nu, N = d.N;
    double* x = calloc((size_t)B * nx, sizeof(double));
    double* r = calloc((size_t)B * nx, sizeof(double));
    double* U = malloc((size_t)B * (N + 1) * nu * sizeof(double));
    int32_t* st = malloc((size_t)B * sizeof(int32_t));
    int32_t* it = calloc((size_t)B, sizeof(int32_t));
    for (int b = 0; b < B; ++b) { x[(size_t)b * nx] = 0.45 * b / (B > 1 ? B - 1 : 1); r[(size_t)b * nx] = 0.5; }
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (rtmpc_qp_solve_host(qp, B, x, r, NULL, 1, 0, NULL, U, st, it) != 0) {
        fprintf(stderr, "rtmpc_qp_solve_host: %s\n", rtmpc_last_error());
        return 1;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    int optimal = 0;
    for (int b = 0; b < B; ++b) optimal += st[b] == RTMPC_OPTIMAL;
    printf("batch %d optimal %d ms %.3f\n", B, optimal, (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6);
    for (int b = 0; b < B; b += (B > 4 ? B / 4 : 1))
        printf("instance %d x1 %.17g status %d steps %d u0 %.17g u_ss %.17g\n", b, x[(size_t)b * nx], st[b], (it[b] >> 12) & 0xFFF,
               U[(size_t)b * (N + 1) * nu], U[(size_t)b * (N + 1) * nu + (size_t)N * nu]);
    rtmpc_qp_destroy(qp);
    return optimal == B ? 0 : 1;
}
