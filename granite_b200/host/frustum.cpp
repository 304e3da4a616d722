#include "frustum.hpp"

#include <cmath>

namespace Granite
{
AABB AABB::transform(const mat_affine &m) const
{
	// per output component: translation + sum over the three source axes of (matrix element *
	// the source bound it selects), accumulated in axis order
	vec3 lo, hi;
	for (int c = 0; c < 3; c++)
	{
		const vec4 &row = m[c];
		float h = row.w, l = row.w;
		for (int k = 0; k < 3; k++)
		{
			const bool positive = row[k] > 0.0f;
			h = h + row[k] * (positive ? maximum[k] : minimum[k]);
			l = l + row[k] * (positive ? minimum[k] : maximum[k]);
		}
		hi[c] = h;
		lo[c] = l;
	}
	return AABB(lo, hi);
}

namespace
{
vec3 unproject(const mat4 &m, float x, float y, float z)
{
	vec4 v = m * vec4(x, y, z, 1.0f);
	return vec3(v.x / v.w, v.y / v.w, v.z / v.w);
}

vec3 face_normal(const vec3 &a0, const vec3 &a1, const vec3 &b0, const vec3 &b1)
{
	return normalize(cross(a0 - a1, b0 - b1));
}
} // namespace

void Frustum::build_planes(const mat4 &inv_view_projection)
{
	const bool infinite_z = inv_view_projection[3][3] == 0.0f;
	const float far_clip_z = infinite_z ? 1e-10f : 0.0f; // FarClipInfiniteClamp
	const vec3 TLN = unproject(inv_view_projection, -1.0f, -1.0f, 1.0f);
	const vec3 BLN = unproject(inv_view_projection, -1.0f, +1.0f, 1.0f);
	const vec3 BLF = unproject(inv_view_projection, -1.0f, +1.0f, far_clip_z);
	const vec3 TRN = unproject(inv_view_projection, +1.0f, -1.0f, 1.0f);
	const vec3 TRF = unproject(inv_view_projection, +1.0f, -1.0f, far_clip_z);
	const vec3 BRN = unproject(inv_view_projection, +1.0f, +1.0f, 1.0f);
	const vec3 BRF = unproject(inv_view_projection, +1.0f, +1.0f, far_clip_z);
	const vec4 center = inv_view_projection * vec4(0.0f, 0.0f, 0.5f, 1.0f);

	const vec3 l = face_normal(BLF, BLN, TLN, BLN);
	const vec3 r = face_normal(TRF, TRN, BRN, TRN);
	const vec3 n = face_normal(BLN, BRN, TRN, BRN);
	const vec3 f = face_normal(TRF, BRF, BLF, BRF);
	const vec3 t = face_normal(TLN, TRN, TRF, TRN);
	const vec3 b = face_normal(BRF, BRN, BLN, BRN);

	planes[0] = vec4(l, -dot(l, BLN));
	planes[1] = vec4(r, -dot(r, TRN));
	planes[2] = vec4(n, -dot(n, BRN));
	planes[3] = infinite_z ? vec4(0.0f) : vec4(f, -dot(f, BRF));
	planes[4] = vec4(t, -dot(t, TRN));
	planes[5] = vec4(b, -dot(b, BRN));
	for (auto &p : planes)
		if (dot(center, p) < 0.0f)
			p = vec4(-p.x, -p.y, -p.z, -p.w);
}

bool Frustum::intersects_fast(const AABB &aabb) const
{
	const vec3 &lo = aabb.get_minimum(), &hi = aabb.get_maximum();
	for (const auto &p : planes)
	{
		// the corner furthest along the plane normal; two horizontal adds: (x + y) + (z + w)
		const float dx = p.x * (p.x > 0.0f ? hi.x : lo.x);
		const float dy = p.y * (p.y > 0.0f ? hi.y : lo.y);
		const float dz = p.z * (p.z > 0.0f ? hi.z : lo.z);
		const float dw = p.w * 1.0f;
		if (std::signbit((dx + dy) + (dz + dw)))
			return false;
	}
	return true;
}
} // namespace Granite
