// lights.cpp -- host-side light preparation (renderer/lights/lights.cpp:55-146, 195-220,
// 330-370 in the reference): falloff range from colour, the packed 48-byte record, the spot
// cone's model matrix and bounding sphere, and view-Z extents for the Z-range binning.
#include "lights.hpp"

#include <limits>

namespace Granite
{
void PositionalLight::set_color(vec3 color_)
{
	color = color_;
	recompute_range();
}

void PositionalLight::set_maximum_range(float range)
{
	cutoff_range = range;
	recompute_range();
}

void PositionalLight::recompute_range()
{
	// distance at which 1/d^2 attenuation of the brightest channel drops below 0.1
	const float target_atten = 0.1f;
	float max_color = max(max(color.x, color.y), color.z);
	set_range(std::sqrt(max_color / target_atten));
}

PositionalFragmentInfo PointLight::get_shader_info(const mat_affine &transform) const
{
	// a scaled node scales the light (uniform scale assumed)
	float scale_factor = transform.get_uniform_scale();
	float max_range = min(falloff_range, cutoff_range) * scale_factor;
	PositionalFragmentInfo info;
	info.color = color * (scale_factor * scale_factor);
	info.spot_scale_bias = {};
	info.position = transform.get_translation();
	info.offset_radius = floatToHalf(vec2(0.0f, max_range));
	info.direction = transform.get_forward();
	info.inv_radius = 1.0f / max_range;
	return info;
}

void SpotLight::set_spot_parameters(float inner_cone_, float outer_cone_)
{
	inner_cone = clamp(inner_cone_, 0.001f, 1.0f);
	outer_cone = clamp(outer_cone_, 0.001f, 1.0f);
	recompute_range();
}

void SpotLight::set_range(float range)
{
	falloff_range = range;
	// tan(outer half-angle): lateral extent of the cone per unit of depth
	xy_range = std::sqrt(1.0f - outer_cone * outer_cone) / outer_cone;
	const float reach = min(falloff_range, cutoff_range);
	const float side = xy_range * reach;
	aabb = AABB(vec3(-side, -side, -reach), vec3(side, side, 0.0f)); // the cone looks down -Z
}

void PointLight::set_range(float range)
{
	falloff_range = range;
	const float reach = min(falloff_range, cutoff_range);
	aabb = AABB(vec3(-reach), vec3(reach));
}

mat_affine SpotLight::build_model_matrix(const mat_affine &transform) const
{
	float max_range = min(falloff_range, cutoff_range);
	return mul_scale(transform, vec3(xy_range * max_range, xy_range * max_range, max_range));
}

PositionalFragmentInfo SpotLight::get_shader_info(const mat_affine &transform) const
{
	float scale_factor = transform.get_uniform_scale();
	float max_range = min(falloff_range, cutoff_range) * scale_factor;

	float spot_scale = 1.0f / max(0.001f, inner_cone - outer_cone);
	float spot_bias = -outer_cone * spot_scale;

	// bounding sphere of the cone: centre on the axis at x = 0.5 * (tan^2 + 1) * R while that
	// is inside the cone's depth, else the base disc's circumsphere
	float tan2 = (1.0f - outer_cone * outer_cone) / (outer_cone * outer_cone);
	float center_distance = ((tan2 + 1.0f) * max_range) * 0.5f;
	float spot_offset, spot_radius;
	if (center_distance < max_range)
	{
		spot_offset = center_distance;
		spot_radius = center_distance;
	}
	else
	{
		spot_offset = max_range;
		spot_radius = std::sqrt(tan2) * max_range;
	}

	PositionalFragmentInfo info;
	info.color = color * (scale_factor * scale_factor);
	info.spot_scale_bias = floatToHalf(vec2(spot_scale, spot_bias));
	info.position = transform.get_translation();
	info.offset_radius = floatToHalf(vec2(spot_offset, spot_radius));
	info.direction = normalize(transform.get_forward());
	info.inv_radius = 1.0f / max_range;
	return info;
}

vec2 point_light_z_range(const RenderContext &context, const vec3 &center, float radius)
{
	auto &pos = context.get_render_parameters().camera_position;
	auto &front = context.get_render_parameters().camera_front;
	float z = dot(center - pos, front);
	return vec2(z - radius, z + radius);
}

vec2 spot_light_z_range(const RenderContext &context, const mat_affine &model)
{
	auto &pos = context.get_render_parameters().camera_position;
	auto &front = context.get_render_parameters().camera_front;
	float lo = std::numeric_limits<float>::infinity();
	float hi = -lo;
	vec3 base_pos = model.get_translation();
	vec3 x_off = model.get_right();
	vec3 y_off = model.get_up();
	vec3 z_base = base_pos + model.get_forward();
	const vec3 hull[5] = {
		base_pos, z_base + x_off + y_off, z_base - x_off + y_off, z_base + x_off - y_off, z_base - x_off - y_off,
	};
	for (auto &p : hull)
	{
		float z = dot(p - pos, front);
		lo = min(z, lo);
		hi = max(z, hi);
	}
	return vec2(lo, hi);
}
} // namespace Granite
