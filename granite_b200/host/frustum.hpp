// frustum.hpp -- the visibility test the light gather applies before the clusterer sees a light
// (renderer/scene.cpp:333-358 gather_positional_lights): world-space AABB of the light's static
// AABB against the six planes of the camera frustum.  Mirrors math/aabb.hpp, math/frustum.hpp and
// the two math/simd.hpp helpers involved (transform_aabb for a mat_affine, frustum_cull), written
// as scalar code with the reference's evaluation order so the kept/culled decision is the same.
#pragma once

#include "math.hpp"

namespace Granite
{
class AABB
{
public:
	AABB() = default;
	AABB(vec3 minimum_, vec3 maximum_) : minimum(minimum_), maximum(maximum_) {}
	const vec3 &get_minimum() const { return minimum; }
	const vec3 &get_maximum() const { return maximum; }
	vec3 get_center() const { return minimum + (maximum - minimum) * 0.5f; } // math/aabb.hpp
	// math/simd.hpp:386-419 SIMD::transform_aabb(output, aabb, mat_affine)
	AABB transform(const mat_affine &m) const;

private:
	vec3 minimum = vec3(0.0f), maximum = vec3(0.0f);
};

class Frustum
{
public:
	// math/frustum.cpp:109-156
	void build_planes(const mat4 &inv_view_projection);
	const vec4 *get_planes() const { return planes; }
	// math/simd.hpp:34-60 SIMD::frustum_cull: true = (conservatively) visible
	bool intersects_fast(const AABB &aabb) const;

private:
	vec4 planes[6];
};
} // namespace Granite
