// TEST INFRASTRUCTURE ONLY — minimal stand-in for the header-only GLM library.
//
// GLM is a build dependency of the reference rasterizer (find_package(glm), reference
// CMakeLists.txt:32) that is not installed in this image and cannot be fetched (no network).
// This file re-implements, from GLM's documented semantics, exactly the subset the reference
// kernels use (vec3 / vec4 / mat3, column-major, `M[c][r]`), so that the reference's own .cu
// files can be compiled UNMODIFIED into oracle/_ref/ as the parity oracle and timing baseline.
//
// Expression shapes follow GLM's documented formulas (e.g. mat3*mat3 result[c][r] =
// a[0][r]*b[c][0] + a[1][r]*b[c][1] + a[2][r]*b[c][2], summed left to right; dot = x*x'+y*y'+z*z')
// because the shape decides how nvcc contracts FMAs and therefore the last-ulp results.
// Nothing in the product (photo-slam_b200/) includes this header.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define GLMS_FN __host__ __device__ __forceinline__
#else
#define GLMS_FN inline
#endif

namespace glm {

struct vec3 {
	float x, y, z;
	GLMS_FN vec3() : x(0), y(0), z(0) {}
	GLMS_FN explicit vec3(float s) : x(s), y(s), z(s) {}
	template <typename A, typename B, typename C>
	GLMS_FN vec3(A a, B b, C c) : x(float(a)), y(float(b)), z(float(c)) {}
	GLMS_FN float& operator[](int i) { return (&x)[i]; }
	GLMS_FN const float& operator[](int i) const { return (&x)[i]; }
	GLMS_FN vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
	GLMS_FN vec3& operator+=(float s) { x += s; y += s; z += s; return *this; }
	GLMS_FN vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
	GLMS_FN vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct vec4 {
	float x, y, z, w;
	GLMS_FN vec4() : x(0), y(0), z(0), w(0) {}
	template <typename A, typename B, typename C, typename D>
	GLMS_FN vec4(A a, B b, C c, D d) : x(float(a)), y(float(b)), z(float(c)), w(float(d)) {}
	GLMS_FN float& operator[](int i) { return (&x)[i]; }
	GLMS_FN const float& operator[](int i) const { return (&x)[i]; }
};

GLMS_FN vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLMS_FN vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLMS_FN vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
GLMS_FN vec3 operator*(float s, const vec3& v) { return vec3(s * v.x, s * v.y, s * v.z); }
GLMS_FN vec3 operator*(const vec3& v, float s) { return vec3(v.x * s, v.y * s, v.z * s); }
GLMS_FN vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
GLMS_FN vec3 operator/(const vec3& v, float s) { return vec3(v.x / s, v.y / s, v.z / s); }

GLMS_FN float dot(const vec3& a, const vec3& b) {
	vec3 t(a.x * b.x, a.y * b.y, a.z * b.z);
	return t.x + t.y + t.z;
}
GLMS_FN float length(const vec3& v) { return sqrtf(dot(v, v)); }
GLMS_FN vec3 max(const vec3& v, float s) {
	return vec3((v.x < s) ? s : v.x, (v.y < s) ? s : v.y, (v.z < s) ? s : v.z);
}

// Column-major 3x3: value[c] is column c; M[c][r].
struct mat3 {
	vec3 value[3];
	GLMS_FN mat3() { value[0] = vec3(1, 0, 0); value[1] = vec3(0, 1, 0); value[2] = vec3(0, 0, 1); }
	GLMS_FN explicit mat3(float s) { value[0] = vec3(s, 0, 0); value[1] = vec3(0, s, 0); value[2] = vec3(0, 0, s); }
	template <typename X1, typename Y1, typename Z1, typename X2, typename Y2, typename Z2, typename X3, typename Y3, typename Z3>
	GLMS_FN mat3(X1 x1, Y1 y1, Z1 z1, X2 x2, Y2 y2, Z2 z2, X3 x3, Y3 y3, Z3 z3) {
		value[0] = vec3(x1, y1, z1); value[1] = vec3(x2, y2, z2); value[2] = vec3(x3, y3, z3);
	}
	GLMS_FN vec3& operator[](int c) { return value[c]; }
	GLMS_FN const vec3& operator[](int c) const { return value[c]; }
};

GLMS_FN mat3 operator*(const mat3& a, const mat3& b) {
	const float a00 = a[0][0], a01 = a[0][1], a02 = a[0][2];
	const float a10 = a[1][0], a11 = a[1][1], a12 = a[1][2];
	const float a20 = a[2][0], a21 = a[2][1], a22 = a[2][2];
	const float b00 = b[0][0], b01 = b[0][1], b02 = b[0][2];
	const float b10 = b[1][0], b11 = b[1][1], b12 = b[1][2];
	const float b20 = b[2][0], b21 = b[2][1], b22 = b[2][2];
	mat3 r(0.0f);
	r[0][0] = a00 * b00 + a10 * b01 + a20 * b02;
	r[0][1] = a01 * b00 + a11 * b01 + a21 * b02;
	r[0][2] = a02 * b00 + a12 * b01 + a22 * b02;
	r[1][0] = a00 * b10 + a10 * b11 + a20 * b12;
	r[1][1] = a01 * b10 + a11 * b11 + a21 * b12;
	r[1][2] = a02 * b10 + a12 * b11 + a22 * b12;
	r[2][0] = a00 * b20 + a10 * b21 + a20 * b22;
	r[2][1] = a01 * b20 + a11 * b21 + a21 * b22;
	r[2][2] = a02 * b20 + a12 * b21 + a22 * b22;
	return r;
}
GLMS_FN mat3 operator*(float s, const mat3& m) {
	mat3 r(0.0f);
	r[0] = m[0] * s; r[1] = m[1] * s; r[2] = m[2] * s;
	return r;
}
GLMS_FN mat3 operator*(const mat3& m, float s) { return s * m; }
GLMS_FN mat3 transpose(const mat3& m) {
	return mat3(m[0][0], m[1][0], m[2][0],
	            m[0][1], m[1][1], m[2][1],
	            m[0][2], m[1][2], m[2][2]);
}

}  // namespace glm
