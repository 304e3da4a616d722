// render_context.hpp -- the per-frame parameter blocks the hot-path builders read:
// RenderParameters (math/render_parameters.hpp:37-59), FrameParameters, LightingParameters
// (directional light), and RenderContext::set_camera (renderer/render_context.cpp:54-87).
#pragma once

#include "frustum.hpp"
#include "math.hpp"

namespace Granite
{
class LightClusterer;

struct RenderParameters
{
	mat4 projection;
	mat4 view;
	mat4 view_projection;
	mat4 inv_projection;
	mat4 inv_view;
	mat4 inv_view_projection;
	vec3 camera_position;
	vec3 camera_front;
	vec3 camera_right;
	vec3 camera_up;
	float z_near = 0.0f;
	float z_far = 0.0f;
};

struct FrameParameters
{
	double frame_time = 0.0;
	double elapsed_time = 0.0;
};

struct DirectionalParameters
{
	vec3 color;
	vec3 direction;
};

struct LightingParameters
{
	DirectionalParameters directional;
	LightClusterer *cluster = nullptr;
};

class RenderContext
{
public:
	void set_camera(const mat4 &projection, const mat4 &view);
	// Takes a parameter block computed elsewhere (e.g. by the application's own camera code)
	// verbatim: the matrices are inputs to this path, not something it derives.
	void set_render_parameters(const RenderParameters &params)
	{
		camera = params;
		frustum.build_planes(camera.inv_view_projection);
	}
	// renderer/render_context.cpp:69, render_context.hpp:108
	const Frustum &get_visibility_frustum() const { return frustum; }
	const RenderParameters &get_render_parameters() const { return camera; }
	void set_frame_parameters(const FrameParameters &frame_) { frame = frame_; }
	const FrameParameters &get_frame_parameters() const { return frame; }
	void set_lighting_parameters(const LightingParameters *lighting_) { lighting = lighting_; }
	const LightingParameters *get_lighting_parameters() const { return lighting; }

private:
	RenderParameters camera;
	Frustum frustum;
	FrameParameters frame;
	const LightingParameters *lighting = nullptr;
};
} // namespace Granite
