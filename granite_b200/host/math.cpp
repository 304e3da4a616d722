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
} // namespace muglm
