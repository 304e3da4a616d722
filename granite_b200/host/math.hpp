// math.hpp -- the small slice of muglm the hot-path builders need (vec2/3/4, mat4, mat_affine),
// written from scratch with muglm's storage layout (column-major mat4, row-major mat_affine:
// math/muglm/muglm.hpp) so POD parameter blocks keep the reference's byte layout.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

namespace muglm
{
struct vec2
{
	float x = 0, y = 0;
	vec2() = default;
	vec2(float x_, float y_) : x(x_), y(y_) {}
	explicit vec2(float v) : x(v), y(v) {}
};

struct vec3
{
	float x = 0, y = 0, z = 0;
	vec3() = default;
	vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
	explicit vec3(float v) : x(v), y(v), z(v) {}
	float &operator[](int i) { return (&x)[i]; }
	const float &operator[](int i) const { return (&x)[i]; }
};

struct vec4
{
	float x = 0, y = 0, z = 0, w = 0;
	vec4() = default;
	vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
	vec4(const vec3 &v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
	explicit vec4(float v) : x(v), y(v), z(v), w(v) {}
	vec3 xyz() const { return vec3(x, y, z); }
	float &operator[](int i) { return (&x)[i]; }
	const float &operator[](int i) const { return (&x)[i]; }
};

struct uvec2
{
	uint32_t x = 0, y = 0;
	uvec2() = default;
	uvec2(uint32_t x_, uint32_t y_) : x(x_), y(y_) {}
};

struct ivec2
{
	int32_t x = 0, y = 0;
	ivec2() = default;
	ivec2(int32_t x_, int32_t y_) : x(x_), y(y_) {}
};

struct u16vec2
{
	uint16_t x = 0, y = 0;
};

inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(const vec3 &a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(const vec3 &a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec4 operator*(const vec4 &a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator+(const vec4 &a, const vec4 &b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline float dot(const vec3 &a, const vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(const vec4 &a, const vec4 &b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float length(const vec3 &a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(const vec3 &a) { return a * (1.0f / length(a)); }
inline vec3 cross(const vec3 &a, const vec3 &b) { return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float min(float a, float b) { return a < b ? a : b; }
inline float max(float a, float b) { return a > b ? a : b; }
inline float clamp(float v, float lo, float hi) { return min(max(v, lo), hi); }

// Column-major, m[col][row] like muglm::mat4.
struct mat4
{
	vec4 vec[4];
	mat4() = default;
	explicit mat4(float d)
	{
		vec[0] = vec4(d, 0, 0, 0);
		vec[1] = vec4(0, d, 0, 0);
		vec[2] = vec4(0, 0, d, 0);
		vec[3] = vec4(0, 0, 0, d);
	}
	mat4(const vec4 &a, const vec4 &b, const vec4 &c, const vec4 &d)
	{
		vec[0] = a; vec[1] = b; vec[2] = c; vec[3] = d;
	}
	vec4 &operator[](int i) { return vec[i]; }
	const vec4 &operator[](int i) const { return vec[i]; }
	const float *data() const { return &vec[0].x; }
	float *data() { return &vec[0].x; }
};
static_assert(sizeof(mat4) == 64, "mat4 layout");

inline vec4 operator*(const mat4 &m, const vec4 &v) { return m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w; }
inline mat4 operator*(const mat4 &a, const mat4 &b) { return mat4(a * b[0], a * b[1], a * b[2], a * b[3]); }

inline mat4 translate(const vec3 &v)
{
	mat4 m(1.0f);
	m[3] = vec4(v, 1.0f);
	return m;
}

inline mat4 scale(const vec3 &v)
{
	mat4 m(1.0f);
	m[0].x = v.x; m[1].y = v.y; m[2].z = v.z;
	return m;
}

// General 4x4 inverse (Gauss-Jordan in double with partial pivoting, rounded once to fp32).
mat4 inverse(const mat4 &m);
// Reverse-Z, Y-flipped perspective; far == InfiniteFarPlane gives the infinite-far form
// (same matrix as muglm::perspective, math/muglm/muglm.cpp:319-345).
mat4 perspective(float fovy, float aspect, float z_near, float z_far);
constexpr float InfiniteFarPlane = 3.402823466e+38f;

// Rotations for the shadow views of the positional lights (math/muglm/muglm.hpp quat: w + (x, y, z)).
struct quat
{
	float w = 1.0f, x = 0.0f, y = 0.0f, z = 0.0f;
	quat() = default;
	quat(float w_, const vec3 &v) : w(w_), x(v.x), y(v.y), z(v.z) {}
	quat(float w_, float x_, float y_, float z_) : w(w_), x(x_), y(y_), z(z_) {}
};
// Shortest rotation taking `from` onto `to` (math/transforms.cpp:122-148), degenerate cases included.
quat rotate_vector(vec3 from, vec3 to);
// Rotation that turns `direction` onto the view axis -Z with whatever roll falls out (math/transforms.cpp:180-183).
quat look_at_arbitrary_up(const vec3 &direction);
// Rotation matrix of a unit quaternion (math/muglm/muglm.cpp:29-62).
mat4 mat4_cast(const quat &q);

// Rows of a 3x4 affine transform, like muglm::mat_affine (math/muglm/muglm.hpp:927-957).
struct mat_affine
{
	vec4 vec[3];
	mat_affine()
	{
		vec[0] = vec4(1, 0, 0, 0);
		vec[1] = vec4(0, 1, 0, 0);
		vec[2] = vec4(0, 0, 1, 0);
	}
	mat_affine(const vec4 &r0, const vec4 &r1, const vec4 &r2)
	{
		vec[0] = r0; vec[1] = r1; vec[2] = r2;
	}
	vec4 &operator[](int i) { return vec[i]; }
	const vec4 &operator[](int i) const { return vec[i]; }
	vec3 get_translation() const { return vec3(vec[0].w, vec[1].w, vec[2].w); }
	vec3 get_right() const { return vec3(vec[0].x, vec[1].x, vec[2].x); }
	vec3 get_up() const { return vec3(vec[0].y, vec[1].y, vec[2].y); }
	vec3 get_forward() const { return vec3(-vec[0].z, -vec[1].z, -vec[2].z); }
	// length of ROW 0's xyz, as muglm does (math/muglm/muglm.cpp:434-437)
	float get_uniform_scale() const { return length(vec3(vec[0].x, vec[0].y, vec[0].z)); }
};
static_assert(sizeof(mat_affine) == 48, "mat_affine layout");

// a * scale(s): scales the basis columns, keeps the translation.
inline mat_affine mul_scale(const mat_affine &a, const vec3 &s)
{
	mat_affine r;
	for (int i = 0; i < 3; i++)
		r[i] = vec4(a[i].x * s.x, a[i].y * s.y, a[i].z * s.z, a[i].w);
	return r;
}

// math/muglm/muglm_impl.hpp:860-907: fp32 -> fp16 rounding half UP on the magnitude (not RNE).
uint16_t floatToHalf(float v);
inline u16vec2 floatToHalf(const vec2 &v)
{
	u16vec2 r;
	r.x = floatToHalf(v.x);
	r.y = floatToHalf(v.y);
	return r;
}
} // namespace muglm

namespace Granite
{
using namespace muglm;
}
