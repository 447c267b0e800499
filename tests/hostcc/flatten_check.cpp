// TEST INFRASTRUCTURE ONLY — host-side check of csrc/compound_flatten.h (the child table of nested CompoundShapes) without CUDA:
// depth-first leaf order, frame chains, the depth limit, and compoundChildWorld == the recursive composition.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../emu/cuda_runtime.h"   // shim of the CUDA built-ins: the REAL common.cuh helpers compile for the host
#include "../../libgdx-jbullet_b200/csrc/common.cuh"
#include "../../libgdx-jbullet_b200/csrc/compound_flatten.h"
using namespace b2c;

static CompoundDirectChild kid(int shape, float angle, float ox, float oy, float oz) {
    CompoundDirectChild c{};
    c.shape = shape;
    const float cs = std::cos(angle), sn = std::sin(angle);
    const float m[9] = {cs, -sn, 0, sn, cs, 0, 0, 0, 1};
    std::memcpy(c.xf12, m, sizeof m);
    c.xf12[9] = ox; c.xf12[10] = oy; c.xf12[11] = oz;
    return c;
}
static Xf xfOf(const CompoundDirectChild& c) {
    Xf t;
    for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) t.m[r][k] = c.xf12[3 * r + k];
    t.o = mk3(c.xf12[9], c.xf12[10], c.xf12[11]);
    return t;
}
int main() {
    // shapes 0..2 are leaves; 10 = {0 @A, 1 @B}; 11 = {10 @C, 2 @D}; 12 = {11 @E, 10 @F, 0 @G}
    std::vector<std::vector<CompoundDirectChild>> direct(20);
    direct[10] = {kid(0, 0.3f, 1, 0, 0), kid(1, -0.2f, 0, 1, 0)};
    direct[11] = {kid(10, 0.7f, 0, 0, 1), kid(2, 0.1f, -1, 0, 0)};
    direct[12] = {kid(11, -0.5f, 0.5f, 0.5f, 0), kid(10, 1.1f, 0, -2, 0), kid(0, 0.f, 3, 0, 0)};
    auto directOf = [&](int s) -> const std::vector<CompoundDirectChild>* { return (s < (int)direct.size() && !direct[s].empty()) ? &direct[s] : nullptr; };
    std::vector<CompoundChildDev> table(7);   // something in front, as in a real table
    int first = 0, n = 0;
    if (!flattenCompound(table, direct[12], directOf, first, n)) { std::printf("FAIL depth\n"); return 1; }
    // depth-first leaves of 12: 11 -> (10 -> 0, 1), 2 ; 10 -> 0, 1 ; 0
    const int expectShape[6] = {0, 1, 2, 0, 1, 0};
    if (n != 6) { std::printf("FAIL leaf count %d\n", n); return 1; }
    for (int i = 0; i < n; i++) if (table[first + i].shape != expectShape[i]) { std::printf("FAIL leaf %d shape %d\n", i, table[first + i].shape); return 1; }
    // frames: 11 (under nothing), 10 under 11, 10 (second occurrence, under nothing) = 3 frame entries in front of the leaves
    if (first != 7 + 3) { std::printf("FAIL frame count %d\n", first - 7); return 1; }
    for (int i = 7; i < first; i++) if (table[i].shape != -1) { std::printf("FAIL frame entry %d is not a frame\n", i); return 1; }
    // world transform of every leaf == the recursion ((org * t1) * t2) * t3, bit for bit
    Xf org = xfOf(kid(0, 0.9f, 10, 20, 30));
    const Xf E = xfOf(direct[12][0]), F = xfOf(direct[12][1]), G = xfOf(direct[12][2]);
    const Xf C = xfOf(direct[11][0]), D = xfOf(direct[11][1]), A = xfOf(direct[10][0]), B = xfOf(direct[10][1]);
    const Xf want[6] = {mulXf(mulXf(mulXf(org, E), C), A), mulXf(mulXf(mulXf(org, E), C), B), mulXf(mulXf(org, E), D),
                        mulXf(mulXf(org, F), A), mulXf(mulXf(org, F), B), mulXf(org, G)};
    for (int i = 0; i < n; i++) {
        const Xf got = compoundChildWorld(org, table.data(), table[first + i]);
        if (std::memcmp(&got, &want[i], sizeof(Xf)) != 0) { std::printf("FAIL leaf %d world transform\n", i); return 1; }
    }
    // depth limit: a chain of 5 nested compounds above a leaf is one frame too many
    std::vector<std::vector<CompoundDirectChild>> deep(10);
    deep[1] = {kid(0, 0, 0, 0, 0)};
    for (int l = 2; l <= 6; l++) deep[l] = {kid(l - 1, 0.1f, 0, 0, 0)};
    auto deepOf = [&](int s) -> const std::vector<CompoundDirectChild>* { return (s >= 1 && s < (int)deep.size() && !deep[s].empty()) ? &deep[s] : nullptr; };
    std::vector<CompoundChildDev> t2;
    if (!flattenCompound(t2, deep[5], deepOf, first, n) || n != 1) { std::printf("FAIL depth 4 must pass\n"); return 1; }
    std::vector<CompoundChildDev> t3;
    if (flattenCompound(t3, deep[6], deepOf, first, n)) { std::printf("FAIL depth 5 must be refused\n"); return 1; }
    std::printf("ALL OK leaves 6 frames 3\n");
    return 0;
}
