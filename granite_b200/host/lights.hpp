// lights.hpp -- positional lights and their 48-byte GPU record, mirroring
// renderer/lights/lights.hpp + light_info.hpp for the members the clusterer path uses.
#pragma once

#include "frustum.hpp"
#include "math.hpp"
#include "render_context.hpp"

namespace Granite
{
// renderer/lights/light_info.hpp:35-44 (== GrbPositionalLight in the C ABI)
struct PositionalFragmentInfo
{
	vec3 color;
	u16vec2 spot_scale_bias;
	vec3 position;
	u16vec2 offset_radius;
	vec3 direction;
	float inv_radius;
};
static_assert(sizeof(PositionalFragmentInfo) == 48, "PositionFragmentInfo is not 48 bytes.");

class PositionalLight
{
public:
	enum class Type
	{
		Spot,
		Point
	};
	explicit PositionalLight(Type type_) : type(type_) {}
	virtual ~PositionalLight() = default;
	Type get_type() const { return type; }
	void set_color(vec3 color_);
	const vec3 &get_color() const { return color; }
	void set_maximum_range(float range);
	float get_maximum_range() const { return min(falloff_range, cutoff_range); }
	// local-space bounds the scene's visibility test transforms by the node transform
	// (renderer/lights/lights.cpp:77-89 spot, :196-201 point; renderer/scene.cpp:1130)
	const AABB &get_static_aabb() const { return aabb; }
	// The light's shadow map (SpotLight / PointLight::set_shadow_info, renderer/lights/lights.cpp:91-95,222-226): device
	// memory the CALLER rendered and owns -- D16_UNORM, resolution^2 texels for a spot light, 6 faces of resolution^2
	// (+X -X +Y -Y +Z -Z) for a point light, resolution = LightClusterer::set_shadow_resolution.  Null: casts no shadow.
	void set_shadow_map(const void *device_d16) { shadow_map = device_d16; }
	const void *get_shadow_map() const { return shadow_map; }

protected:
	vec3 color = vec3(1.0f);
	float falloff_range = 1.0f;
	float cutoff_range = 1e10f;
	AABB aabb;
	void recompute_range();
	virtual void set_range(float range) = 0;

private:
	Type type;
	const void *shadow_map = nullptr;
};

class PointLight : public PositionalLight
{
public:
	PointLight() : PositionalLight(Type::Point) {}
	PositionalFragmentInfo get_shader_info(const mat_affine &transform) const;

private:
	void set_range(float range) override;
};

class SpotLight : public PositionalLight
{
public:
	SpotLight() : PositionalLight(Type::Spot) {}
	void set_spot_parameters(float inner_cone, float outer_cone);
	PositionalFragmentInfo get_shader_info(const mat_affine &transform) const;
	mat_affine build_model_matrix(const mat_affine &transform) const;
	float get_xy_range() const { return xy_range; }

private:
	float inner_cone = 0.4f;
	float outer_cone = 0.45f;
	float xy_range = 0.0f;
	void set_range(float range) override;
};

vec2 point_light_z_range(const RenderContext &context, const vec3 &center, float radius);
vec2 spot_light_z_range(const RenderContext &context, const mat_affine &model);
} // namespace Granite
