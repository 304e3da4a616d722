#include "math.hpp"

#include <algorithm>

namespace muglm
{
mat4 inverse(const mat4 &m)
{
	// Augmented [A | I] elimination in double; A is column-major so a(r, c) = m[c][r].
	double a[4][8];
	for (int r = 0; r < 4; r++)
		for (int c = 0; c < 4; c++)
		{
			a[r][c] = (double)m[c][r];
			a[r][4 + c] = r == c ? 1.0 : 0.0;
		}
	for (int col = 0; col < 4; col++)
	{
		int pivot = col;
		for (int r = col + 1; r < 4; r++)
			if (std::fabs(a[r][col]) > std::fabs(a[pivot][col]))
				pivot = r;
		if (pivot != col)
			for (int c = 0; c < 8; c++)
				std::swap(a[pivot][c], a[col][c]);
		double inv = 1.0 / a[col][col];
		for (int c = 0; c < 8; c++)
			a[col][c] *= inv;
		for (int r = 0; r < 4; r++)
		{
			if (r == col)
				continue;
			double f = a[r][col];
			if (f == 0.0)
				continue;
			for (int c = 0; c < 8; c++)
				a[r][c] -= f * a[col][c];
		}
	}
	mat4 out;
	for (int r = 0; r < 4; r++)
		for (int c = 0; c < 4; c++)
			out[c][r] = (float)a[r][4 + c] + 0.0f; // + 0: no negative zeros
	return out;
}

mat4 perspective(float fovy, float aspect, float z_near, float z_far)
{
	const float t = std::tan(fovy / 2.0f);
	mat4 p(0.0f);
	p[0][0] = 1.0f / (aspect * t);
	p[1][1] = -(1.0f / t); // Y flip folded in
	p[2][3] = -1.0f;
	if (z_far == InfiniteFarPlane)
		p[3][2] = z_near; // reverse-Z, infinite far
	else
	{
		p[2][2] = -1.0f - z_far / (z_near - z_far);
		p[3][2] = -(z_far * z_near) / (z_near - z_far);
	}
	return p;
}

uint16_t floatToHalf(float v)
{
	uint32_t u;
	std::memcpy(&u, &v, 4);
	const uint32_t sign = (u >> 16) & 0x8000u;
	const uint32_t mag = u & 0x7fffffffu;
	if (mag >= 0x7f800000u)
	{
		uint32_t payload = (mag & 0x7fffffu) >> 13;
		if ((mag & 0x7fffffu) != 0 && payload == 0)
			payload = 1; // keep NaN a NaN
		return (uint16_t)(sign | 0x7c00u | payload);
	}
	const int e = (int)(mag >> 23) - 112; // rebias 127 -> 15
	if (e <= 0)
	{
		if (e < -10)
			return (uint16_t)sign; // underflows to zero
		// denormal half: shift in the hidden bit, add half an output ulp, truncate
		uint32_t m = ((mag & 0x7fffffu) | 0x800000u) >> (1 - e);
		return (uint16_t)(sign | ((m + 0x1000u) >> 13));
	}
	// normal: "round half up on the magnitude" is +half-ulp then truncate; a mantissa carry
	// rolls into the exponent field on its own
	uint32_t h = (((uint32_t)e << 23) | (mag & 0x7fffffu)) + 0x1000u;
	h >>= 13;
	if (h >= 0x7c00u)
		h = 0x7c00u;
	return (uint16_t)(sign | h);
}

quat rotate_vector(vec3 from, vec3 to)
{
	from = normalize(from);
	to = normalize(to);
	const float cos_angle = dot(from, to);
	if (std::fabs(cos_angle) > 0.9999f)
	{
		if (cos_angle > 0.9999f)
			return quat(1.0f, 0.0f, 0.0f, 0.0f);
		// opposite vectors: half a turn about any axis perpendicular to `from`
		vec3 axis = cross(vec3(1.0f, 0.0f, 0.0f), from);
		if (dot(axis, axis) > 0.001f)
			axis = normalize(axis);
		else
			axis = normalize(cross(vec3(0.0f, 1.0f, 0.0f), from));
		return quat(0.0f, axis);
	}
	const vec3 axis = normalize(cross(from, to));
	const vec3 half_vector = normalize(from + to);
	const float cos_half = clamp(dot(half_vector, from), 0.0f, 1.0f);
	const float sin_half = std::sqrt(1.0f - cos_half * cos_half);
	return quat(cos_half, axis * sin_half);
}

quat look_at_arbitrary_up(const vec3 &direction)
{
	return rotate_vector(normalize(direction), vec3(0.0f, 0.0f, -1.0f));
}

mat4 mat4_cast(const quat &q)
{
	const float xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
	const float xz = q.x * q.z, xy = q.x * q.y, yz = q.y * q.z;
	const float wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z;
	mat4 m(1.0f);
	m[0] = vec4(1.0f - 2.0f * (yy + zz), 2.0f * (xy + wz), 2.0f * (xz - wy), 0.0f);
	m[1] = vec4(2.0f * (xy - wz), 1.0f - 2.0f * (xx + zz), 2.0f * (yz + wx), 0.0f);
	m[2] = vec4(2.0f * (xz + wy), 2.0f * (yz - wx), 1.0f - 2.0f * (xx + yy), 0.0f);
	return m;
}
} // namespace muglm
