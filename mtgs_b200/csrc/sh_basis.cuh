// Real spherical-harmonics basis (degrees 0..4) and its gradient, shared by sh.cu and arena.cu.
// Constants: the published "efficient SH evaluation" polynomials; upstream: gsplat v1.4.0 compute_sh.
#pragma once

template <int DEG>
__device__ __forceinline__ void sh_basis(float x, float y, float z, float *B) {
    B[0] = 0.2820947917738781f;
    if (DEG < 1) return;
    B[1] = -0.48860251190292f * y;
    B[2] = 0.48860251190292f * z;
    B[3] = -0.48860251190292f * x;
    if (DEG < 2) return;
    const float z2 = z * z;
    const float fTmp0B = -1.092548430592079f * z;
    const float fC1 = x * x - y * y, fS1 = 2.f * x * y;
    B[6] = 0.9461746957575601f * z2 - 0.3153915652525201f;
    B[7] = fTmp0B * x;
    B[5] = fTmp0B * y;
    B[8] = 0.5462742152960395f * fC1;
    B[4] = 0.5462742152960395f * fS1;
    if (DEG < 3) return;
    const float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
    const float fTmp1B = 1.445305721320277f * z;
    const float fC2 = x * fC1 - y * fS1, fS2 = x * fS1 + y * fC1;
    B[12] = z * (1.865881662950577f * z2 - 1.119528997770346f);
    B[13] = fTmp0C * x;
    B[11] = fTmp0C * y;
    B[14] = fTmp1B * fC1;
    B[10] = fTmp1B * fS1;
    B[15] = -0.5900435899266435f * fC2;
    B[9] = -0.5900435899266435f * fS2;
    if (DEG < 4) return;
    const float fTmp0D = z * (-4.683325804901025f * z2 + 2.007139630671868f);
    const float fTmp1C = 3.31161143515146f * z2 - 0.47308734787878f;
    const float fTmp2B = -1.770130769779931f * z;
    const float fC3 = x * fC2 - y * fS2, fS3 = x * fS2 + y * fC2;
    B[20] = 1.984313483298443f * z * B[12] + -1.006230589874905f * B[6];
    B[21] = fTmp0D * x;
    B[19] = fTmp0D * y;
    B[22] = fTmp1C * fC1;
    B[18] = fTmp1C * fS1;
    B[23] = fTmp2B * fC2;
    B[17] = fTmp2B * fS2;
    B[24] = 0.6258357354491763f * fC3;
    B[16] = 0.6258357354491763f * fS3;
}

// d(basis_k)/d(x,y,z) for the unit direction; dB[k] = (dx, dy, dz)
template <int DEG>
__device__ __forceinline__ void sh_basis_grad(float x, float y, float z, float3 *dB) {
    dB[0] = make_float3(0.f, 0.f, 0.f);
    if (DEG < 1) return;
    dB[1] = make_float3(0.f, -0.48860251190292f, 0.f);
    dB[2] = make_float3(0.f, 0.f, 0.48860251190292f);
    dB[3] = make_float3(-0.48860251190292f, 0.f, 0.f);
    if (DEG < 2) return;
    const float z2 = z * z;
    const float c0B = -1.092548430592079f, c1 = 0.5462742152960395f;
    const float fC1 = x * x - y * y, fS1 = 2.f * x * y;
    // fC1: (2x, -2y, 0)   fS1: (2y, 2x, 0)
    dB[4] = make_float3(c1 * 2.f * y, c1 * 2.f * x, 0.f);
    dB[5] = make_float3(0.f, c0B * z, c0B * y);
    dB[6] = make_float3(0.f, 0.f, 2.f * 0.9461746957575601f * z);
    dB[7] = make_float3(c0B * z, 0.f, c0B * x);
    dB[8] = make_float3(c1 * 2.f * x, -c1 * 2.f * y, 0.f);
    if (DEG < 3) return;
    const float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
    const float dTmp0C = -2.f * 2.285228997322329f * z;
    const float c1B = 1.445305721320277f, c3 = -0.5900435899266435f;
    const float fC2 = x * fC1 - y * fS1, fS2 = x * fS1 + y * fC1;
    // fC2 = x^3 - 3 x y^2 : (3(x^2-y^2), -6xy, 0) = (3 fC1, -3 fS1, 0)
    // fS2 = 3 x^2 y - y^3 : (6xy, 3(x^2-y^2), 0) = (3 fS1, 3 fC1, 0)
    dB[9] = make_float3(c3 * 3.f * fS1, c3 * 3.f * fC1, 0.f);
    dB[10] = make_float3(c1B * z * 2.f * y, c1B * z * 2.f * x, c1B * fS1);
    dB[11] = make_float3(0.f, fTmp0C, dTmp0C * y);
    dB[12] = make_float3(0.f, 0.f, 3.f * 1.865881662950577f * z2 - 1.119528997770346f);
    dB[13] = make_float3(fTmp0C, 0.f, dTmp0C * x);
    dB[14] = make_float3(c1B * z * 2.f * x, -c1B * z * 2.f * y, c1B * fC1);
    dB[15] = make_float3(c3 * 3.f * fC1, -c3 * 3.f * fS1, 0.f);
    if (DEG < 4) return;
    const float fTmp0D = z * (-4.683325804901025f * z2 + 2.007139630671868f);
    const float dTmp0D = -3.f * 4.683325804901025f * z2 + 2.007139630671868f;
    const float fTmp1C = 3.31161143515146f * z2 - 0.47308734787878f;
    const float dTmp1C = 2.f * 3.31161143515146f * z;
    const float c2B = -1.770130769779931f, c4 = 0.6258357354491763f;
    // fC3 = x fC2 - y fS2 : d/dx = fC2 + x*3fC1 - y*3fS1 = 4 fC2 ; d/dy = -3x fS1 - fS2 - 3y fC1 = -4 fS2
    // fS3 = x fS2 + y fC2 : d/dx = fS2 + 3x fS1 + 3y fC1 = 4 fS2 ; d/dy = 3x fC1 + fC2 - 3y fS1 = 4 fC2
    const float B12 = z * (1.865881662950577f * z2 - 1.119528997770346f);
    const float dB12 = 3.f * 1.865881662950577f * z2 - 1.119528997770346f;
    const float dB6 = 2.f * 0.9461746957575601f * z;
    dB[16] = make_float3(c4 * 4.f * fS2, c4 * 4.f * fC2, 0.f);
    dB[17] = make_float3(c2B * z * 3.f * fS1, c2B * z * 3.f * fC1, c2B * fS2);
    dB[18] = make_float3(fTmp1C * 2.f * y, fTmp1C * 2.f * x, dTmp1C * fS1);
    dB[19] = make_float3(0.f, fTmp0D, dTmp0D * y);
    dB[20] = make_float3(0.f, 0.f, 1.984313483298443f * (B12 + z * dB12) - 1.006230589874905f * dB6);
    dB[21] = make_float3(fTmp0D, 0.f, dTmp0D * x);
    dB[22] = make_float3(fTmp1C * 2.f * x, -fTmp1C * 2.f * y, dTmp1C * fC1);
    dB[23] = make_float3(c2B * z * 3.f * fC1, -c2B * z * 3.f * fS1, c2B * fC2);
    dB[24] = make_float3(c4 * 4.f * fC2, -c4 * 4.f * fS2, 0.f);
}

