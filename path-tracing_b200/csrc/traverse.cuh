// traverse.cuh — software traceRayEXT: BVH4 traversal + watertight ray/triangle test
// (Woop, Benthin, Wald, JCGT 2013) + the two any-hit shaders, for sm_100a.
//
// Replaces the fixed-function TLAS->BLAS traversal behind PT/Shaders/raygen.rgen:31,68.
// Closest-hit contract: nearest accepted hit with tmin < t < tmax; equal t resolved towards the
// smaller flattened triangle id (deterministic, BVH-independent).  Non-opaque candidates run the
// alpha test of anyhit.rahit:36-65 (closest) / occlusionAnyhit.rahit:35-54 (shadow).
#pragma once
#include "scene.cuh"

namespace pt
{

#define PT_STACK_SIZE 64

struct Hit
{
    uint32_t tri; // leaf-order triangle index, 0xffffffff = miss
    float t, b1, b2;
};

// decal record of anyhit.rahit:54-62
struct Decal
{
    float dist; // -1 = none
    float r, g, b, a;
};

struct TraversalStats
{
    uint32_t boxTests, triTests, alphaTests;
};

struct RaySetup
{
    vec3 org;
    float idx, idy, idz; // 1 / dir
    float Sx, Sy, Sz;
    int kx, ky, kz;
};

PT_DEV float comp(vec3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

PT_DEV RaySetup setupRay(vec3 o, vec3 d)
{
    RaySetup r;
    r.org = o;
    r.idx = 1.0f / d.x;
    r.idy = 1.0f / d.y;
    r.idz = 1.0f / d.z;
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    r.kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    r.kx = r.kz == 2 ? 0 : r.kz + 1;
    r.ky = r.kx == 2 ? 0 : r.kx + 1;
    const float dz = comp(d, r.kz);
    if (dz < 0.0f)
    {
        const int t = r.kx;
        r.kx = r.ky;
        r.ky = t;
    }
    r.Sx = comp(d, r.kx) / dz;
    r.Sy = comp(d, r.ky) / dz;
    r.Sz = 1.0f / dz;
    return r;
}

// Returns true with (t, b1, b2) if the line hits the triangle; the caller checks the t range.
PT_DEV bool intersectTriangle(const RaySetup &r, vec3 p0, vec3 p1, vec3 p2, float &t, float &b1, float &b2)
{
    const vec3 A = p0 - r.org, B = p1 - r.org, C = p2 - r.org;
    const float Akz = comp(A, r.kz), Bkz = comp(B, r.kz), Ckz = comp(C, r.kz);
    const float Ax = comp(A, r.kx) - r.Sx * Akz, Ay = comp(A, r.ky) - r.Sy * Akz;
    const float Bx = comp(B, r.kx) - r.Sx * Bkz, By = comp(B, r.ky) - r.Sy * Bkz;
    const float Cx = comp(C, r.kx) - r.Sx * Ckz, Cy = comp(C, r.ky) - r.Sy * Ckz;
    // The edge functions must NOT be contracted into FMAs: watertightness relies on the two
    // triangles sharing an edge computing exactly opposite values (a*b - c*d with both products
    // rounded), which fma(a, b, -(c*d)) breaks — it also makes zero-area triangles "hit".
    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f)
    {
        U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
        V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
        W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f))
        return false;
    const float det = U + V + W;
    if (det == 0.0f)
        return false;
    const float T = U * (r.Sz * Akz) + V * (r.Sz * Bkz) + W * (r.Sz * Ckz);
    const float rcpDet = 1.0f / det;
    t = T * rcpDet;
    b1 = V * rcpDet;
    b2 = W * rcpDet;
    return true;
}

// colour texture x colour factor at a candidate hit (anyhit.rahit:36-51, occlusionAnyhit.rahit:35-50)
PT_DEV float4 anyHitColor(const DeviceScene &s, uint32_t tri, uint32_t materialId, float b1, float b2)
{
    const TriShade &ts = s.triShade[tri];
    const float4 a6 = __ldg(&ts.a[6]), a7 = __ldg(&ts.a[7]), a8 = __ldg(&ts.a[8]);
    const float b0 = 1.0f - b1 - b2;
    const float u = a6.w * b0 + a7.y * b1 + a7.w * b2;
    const float v = a7.x * b0 + a7.z * b1 + a8.x * b2;
    const uint32_t type = materialId & 0xffu, index = materialId >> 8;
    if (type > 2)
        return make_float4(1.0f, 0.0f, 0.0f, 1.0f); // getColorFactor default, texture 0 is white
    const MaterialRaw *m = (type == 0 ? s.materials[0] : type == 1 ? s.materials[1] : s.materials[2]) + index;
    const float4 factor = __ldg(&m->q[1]); // vec4 Color sits at byte 16 in all three structs
    // ColorIdx: MR byte 80 (q[5].x); SG / Phong byte 76 (q[4].w)
    const uint32_t colorIdx =
        type == 0 ? __float_as_uint(__ldg(&m->q[5]).x) : __float_as_uint(__ldg(&m->q[4]).w);
    const float4 c = textureLod0(s, s.textures[colorIdx], u, v);
    return make_float4(c.x * factor.x, c.y * factor.y, c.z * factor.z, c.w * factor.w);
}

// One BVH4 node: tests the four child boxes, returns entry distances (INF if missed / empty).
PT_DEV void intersectNode(const BvhNode *__restrict__ node, const RaySetup &r, float tmin, float tmax, float d[4],
                          int c[4])
{
    const float4 lox = __ldg(&node->lox), loy = __ldg(&node->loy), loz = __ldg(&node->loz);
    const float4 hix = __ldg(&node->hix), hiy = __ldg(&node->hiy), hiz = __ldg(&node->hiz);
    const int4 ch = __ldg(&node->child);
    const float lx[4] = { lox.x, lox.y, lox.z, lox.w }, ly[4] = { loy.x, loy.y, loy.z, loy.w };
    const float lz[4] = { loz.x, loz.y, loz.z, loz.w }, hx[4] = { hix.x, hix.y, hix.z, hix.w };
    const float hy[4] = { hiy.x, hiy.y, hiy.z, hiy.w }, hz[4] = { hiz.x, hiz.y, hiz.z, hiz.w };
    c[0] = ch.x, c[1] = ch.y, c[2] = ch.z, c[3] = ch.w;
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        const float ax = (lx[i] - r.org.x) * r.idx, bx = (hx[i] - r.org.x) * r.idx;
        const float ay = (ly[i] - r.org.y) * r.idy, by = (hy[i] - r.org.y) * r.idy;
        const float az = (lz[i] - r.org.z) * r.idz, bz = (hz[i] - r.org.z) * r.idz;
        // fminf/fmaxf drop NaNs (0 * inf on flat boxes)
        const float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
        const float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
        // conservative (Ize 2013): never cull a box the triangle test could still hit
        d[i] = (c[i] != PT_CHILD_EMPTY && t0 <= t1 * 1.0000004f) ? t0 : INFINITY;
    }
}

#define PT_CSWAP(i, j)                                                                                                \
    if (d[j] < d[i])                                                                                                  \
    {                                                                                                                 \
        const float td = d[i];                                                                                        \
        d[i] = d[j];                                                                                                  \
        d[j] = td;                                                                                                    \
        const int tc = c[i];                                                                                          \
        c[i] = c[j];                                                                                                  \
        c[j] = tc;                                                                                                    \
    }

// CLOSEST = true : nearest hit (+ decal record if ALPHA)
// CLOSEST = false: any hit in (tmin, tmax) with alpha >= 1 -> hit.tri != miss
template <bool CLOSEST, bool ALPHA, bool STATS>
PT_DEV void traverse(const DeviceScene &s, vec3 org, vec3 dir, float tmin, float tmax, Hit &hit, Decal &decal,
                     TraversalStats &st)
{
    hit.tri = 0xffffffffu;
    hit.t = tmax;
    hit.b1 = hit.b2 = 0.0f;
    uint32_t bestFlat = 0xffffffffu;
    if (ALPHA)
        decal.dist = -1.0f;
    if (s.triCount == 0)
        return;
    // A ray with a non-finite component or a zero direction cannot hit anything (every triangle
    // test evaluates to NaN), but NaN also defeats box culling, so it would walk the WHOLE tree.
    // Such rays exist by design: refract() returns 0 on total internal reflection and
    // normalize(0) is NaN; the sample is then restarted (Q7).  Miss immediately.
    {
        const float sum = org.x + org.y + org.z + dir.x + dir.y + dir.z;
        if (!isfinite(sum) || (dir.x == 0.0f && dir.y == 0.0f && dir.z == 0.0f))
            return;
    }
    const RaySetup r = setupRay(org, dir);

    int stackNode[PT_STACK_SIZE];
    float stackDist[PT_STACK_SIZE];
    int sp = 0;
    int cur = 0; // root is always an internal node
    float best = tmax;

    for (;;)
    {
        if (cur >= 0)
        {
            float d[4];
            int c[4];
            intersectNode(s.nodes + cur, r, tmin, best, d, c);
            if (STATS)
                st.boxTests += 4;
            // sorting network, ascending by distance (missed children carry INF)
            PT_CSWAP(0, 1)
            PT_CSWAP(2, 3)
            PT_CSWAP(0, 2)
            PT_CSWAP(1, 3)
            PT_CSWAP(1, 2)
            if (d[0] == INFINITY)
            {
                // nothing hit: pop
                cur = PT_CHILD_EMPTY;
            }
            else
            {
                cur = c[0];
                // push the rest, farthest first
                if (d[3] != INFINITY && sp < PT_STACK_SIZE)
                {
                    stackNode[sp] = c[3];
                    stackDist[sp++] = d[3];
                }
                if (d[2] != INFINITY && sp < PT_STACK_SIZE)
                {
                    stackNode[sp] = c[2];
                    stackDist[sp++] = d[2];
                }
                if (d[1] != INFINITY && sp < PT_STACK_SIZE)
                {
                    stackNode[sp] = c[1];
                    stackDist[sp++] = d[1];
                }
                continue;
            }
        }
        else
        {
            // leaf
            const uint32_t code = (uint32_t)~cur;
            const uint32_t first = code >> 2, count = (code & 3u) + 1;
            for (uint32_t i = 0; i < count; i++)
            {
                const uint32_t tri = first + i;
                const float4 q0 = __ldg(s.triPos + 3 * (size_t)tri);
                const float4 q1 = __ldg(s.triPos + 3 * (size_t)tri + 1);
                const float4 q2 = __ldg(s.triPos + 3 * (size_t)tri + 2);
                if (STATS)
                    st.triTests++;
                float t, b1, b2;
                if (!intersectTriangle(r, V3(q0), V3(q1), V3(q2), t, b1, b2))
                    continue;
                if (!(t > tmin))
                    continue;
                const uint32_t flat = __float_as_uint(q0.w);
                if (CLOSEST)
                {
                    if (!(t < best || (t == best && flat < bestFlat)))
                        continue;
                }
                else if (!(t < tmax))
                    continue;
                if (ALPHA && !(__float_as_uint(q1.w) & PT_TRI_FLAG_OPAQUE))
                {
                    if (STATS)
                        st.alphaTests++;
                    const float4 color = anyHitColor(s, tri, __float_as_uint(q2.w), b1, b2);
                    if (CLOSEST)
                    {
                        if (color.w < 0.5f)
                        {
                            if (decal.dist == -1.0f || t < decal.dist)
                            {
                                decal.r = color.x, decal.g = color.y, decal.b = color.z, decal.a = color.w;
                                decal.dist = t;
                            }
                            continue; // ignoreIntersectionEXT
                        }
                    }
                    else if (color.w < 1.0f)
                        continue;
                }
                hit.tri = tri;
                hit.t = t;
                hit.b1 = b1;
                hit.b2 = b2;
                if (!CLOSEST)
                    return; // gl_RayFlagsTerminateOnFirstHitEXT
                best = t;
                bestFlat = flat;
            }
        }
        // pop, skipping sub-trees that start beyond the current best (ties are kept)
        for (;;)
        {
            if (sp == 0)
                return;
            --sp;
            cur = stackNode[sp];
            if (!CLOSEST || stackDist[sp] <= best)
                break;
        }
    }
}

} // namespace pt
