// Headless stand-in for DirectXTK's <SimpleMath.h>.  TEST INFRASTRUCTURE ONLY.
//
// The reference's simulation sources (src/Sim/*.cpp) are compiled, unmodified, against
// this header to produce the parity oracle (oracle/_ref).  DirectXTK's real SimpleMath
// is a wrapper over DirectXMath, which ships with the Windows SDK and is not part of
// the reference tree, so the arithmetic each member performs is restated here as plain
// scalar fp32 code following DirectXMath's SSE2 (non-FMA) code path:
//
//   * dot / length^2            = (x*x + y*y) + z*z            (three products, two adds)
//   * Length                    = sqrtf(dot)
//   * Normalize                 = v / sqrtf(dot)  (true division; zero vector -> 0;
//                                 infinite length -> NaN)      [XMVector3Normalize]
//   * operator/=(float S)       = v * (1.f / S)                [SimpleMath.inl:778-786]
//   * operator/(Vector3, float) = true division per component  [DirectXMath operator/]
//   * Lerp(a, b, t)             = (b - a) * t + a              [XMVectorLerp]
//   * Cross                     = (y1*z2 - z1*y2, z1*x2 - x1*z2, x1*y2 - y1*x2)
//   * Transform(v, M)           = ((z*r2 + r3) + y*r1) + x*r0, then divide by w
//                                 [XMVector3TransformCoord]
//   * CreateFromYawPitchRoll    = quaternion from half angles -> rotation matrix
//                                 [XMMatrixRotationRollPitchYaw, DirectXMath 3.14], with
//                                 libm sinf/cosf in place of DirectXMath's polynomials.
//
// Build with -ffp-contract=off and without -ffast-math so the compiler does not fuse or
// reassociate any of the above.  This header is the *definition* of those operations for
// every parity claim made in this repository ("Oracle"); bit-parity with
// a Windows build of the reference is not claimed.
#pragma once

#include <cmath>
#include <cstddef>
#include <limits>

namespace DirectX
{
    const float XM_PI = 3.141592654f;
    const float XM_2PI = 6.283185307f;

    namespace SimpleMath
    {
        struct Matrix;

        struct Vector2
        {
            float x, y;
            Vector2() : x(0.f), y(0.f) {}
            Vector2(float ix, float iy) : x(ix), y(iy) {}
        };

        struct Vector3
        {
            float x, y, z;

            Vector3() : x(0.f), y(0.f), z(0.f) {}
            Vector3(float ix, float iy, float iz) : x(ix), y(iy), z(iz) {}

            bool operator==(const Vector3& v) const { return x == v.x && y == v.y && z == v.z; }
            bool operator!=(const Vector3& v) const { return !(*this == v); }

            Vector3& operator+=(const Vector3& v) { x += v.x; y += v.y; z += v.z; return *this; }
            Vector3& operator-=(const Vector3& v) { x -= v.x; y -= v.y; z -= v.z; return *this; }
            Vector3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
            Vector3& operator/=(float s) { const float r = 1.f / s; x *= r; y *= r; z *= r; return *this; }
            Vector3 operator-() const { return Vector3(-x, -y, -z); }

            float LengthSquared() const { return (x * x + y * y) + z * z; }
            float Length() const { return sqrtf(LengthSquared()); }

            Vector3 Cross(const Vector3& v) const
            {
                return Vector3(y * v.z - z * v.y, z * v.x - x * v.z, x * v.y - y * v.x);
            }

            void Normalize()
            {
                const float len = sqrtf(LengthSquared());
                if (len == 0.f) { x = y = z = 0.f; return; }
                if (std::isinf(len))
                {
                    x = y = z = std::numeric_limits<float>::quiet_NaN();
                    return;
                }
                x = x / len; y = y / len; z = z / len;
            }

            static float DistanceSquared(const Vector3& a, const Vector3& b)
            {
                const float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
                return (dx * dx + dy * dy) + dz * dz;
            }

            static Vector3 Lerp(const Vector3& a, const Vector3& b, float t)
            {
                return Vector3((b.x - a.x) * t + a.x, (b.y - a.y) * t + a.y, (b.z - a.z) * t + a.z);
            }

            static inline Vector3 Transform(const Vector3& v, const Matrix& m);

            static const Vector3 Zero;
        };

        inline Vector3 operator+(const Vector3& a, const Vector3& b) { return Vector3(a.x + b.x, a.y + b.y, a.z + b.z); }
        inline Vector3 operator-(const Vector3& a, const Vector3& b) { return Vector3(a.x - b.x, a.y - b.y, a.z - b.z); }
        inline Vector3 operator*(const Vector3& a, float s) { return Vector3(a.x * s, a.y * s, a.z * s); }
        inline Vector3 operator*(float s, const Vector3& a) { return Vector3(a.x * s, a.y * s, a.z * s); }
        inline Vector3 operator/(const Vector3& a, float s) { return Vector3(a.x / s, a.y / s, a.z / s); }

        struct Color
        {
            float x, y, z, w;
            Color() : x(0.f), y(0.f), z(0.f), w(1.f) {}
            Color(float r, float g, float b) : x(r), y(g), z(b), w(1.f) {}
            Color(float r, float g, float b, float a) : x(r), y(g), z(b), w(a) {}
            float R() const { return x; }
            float G() const { return y; }
            float B() const { return z; }
            float A() const { return w; }
        };

        struct Matrix
        {
            float m[4][4];

            Matrix()
            {
                for (int r = 0; r < 4; ++r)
                    for (int c = 0; c < 4; ++c)
                        m[r][c] = (r == c) ? 1.f : 0.f;
            }

            static Matrix CreateScale(float s)
            {
                Matrix R;
                R.m[0][0] = R.m[1][1] = R.m[2][2] = s;
                return R;
            }

            static Matrix CreateTranslation(const Vector3& p)
            {
                Matrix R;
                R.m[3][0] = p.x; R.m[3][1] = p.y; R.m[3][2] = p.z;
                return R;
            }

            // Quaternion route of XMMatrixRotationRollPitchYaw(pitch, yaw, roll).
            static Matrix CreateFromYawPitchRoll(float yaw, float pitch, float roll)
            {
                const float hp = pitch * 0.5f, hy = yaw * 0.5f, hr = roll * 0.5f;
                const float sp = sinf(hp), cp = cosf(hp);
                const float sy = sinf(hy), cy = cosf(hy);
                const float sr = sinf(hr), cr = cosf(hr);

                const float qx = (cr * sp) * cy + (sr * cp) * sy;
                const float qy = (cr * cp) * sy - (sr * sp) * cy;
                const float qz = (sr * cp) * cy - (cr * sp) * sy;
                const float qw = (cr * cp) * cy + (sr * sp) * sy;

                const float xx = qx * qx, yy = qy * qy, zz = qz * qz;
                const float xy = qx * qy, xz = qx * qz, yz = qy * qz;
                const float wx = qw * qx, wy = qw * qy, wz = qw * qz;

                Matrix R;
                R.m[0][0] = 1.f - 2.f * (yy + zz);
                R.m[0][1] = 2.f * (xy + wz);
                R.m[0][2] = 2.f * (xz - wy);
                R.m[1][0] = 2.f * (xy - wz);
                R.m[1][1] = 1.f - 2.f * (xx + zz);
                R.m[1][2] = 2.f * (yz + wx);
                R.m[2][0] = 2.f * (xz + wy);
                R.m[2][1] = 2.f * (yz - wx);
                R.m[2][2] = 1.f - 2.f * (xx + yy);
                return R;
            }

            Matrix operator*(const Matrix& b) const
            {
                Matrix R;
                for (int r = 0; r < 4; ++r)
                    for (int c = 0; c < 4; ++c)
                    {
                        float s = m[r][0] * b.m[0][c];
                        s = s + m[r][1] * b.m[1][c];
                        s = s + m[r][2] * b.m[2][c];
                        s = s + m[r][3] * b.m[3][c];
                        R.m[r][c] = s;
                    }
                return R;
            }
        };

        inline Vector3 Vector3::Transform(const Vector3& v, const Matrix& M)
        {
            float r[4];
            for (int c = 0; c < 4; ++c)
            {
                float s = v.z * M.m[2][c] + M.m[3][c];
                s = v.y * M.m[1][c] + s;
                s = v.x * M.m[0][c] + s;
                r[c] = s;
            }
            return Vector3(r[0] / r[3], r[1] / r[3], r[2] / r[3]);
        }
    }
}
