// render_context.cpp -- RenderContext::set_camera (renderer/render_context.cpp:54-87): derives
// the matrices and camera vectors the hot path reads from (projection, view).
#include "render_context.hpp"

namespace Granite
{
void RenderContext::set_camera(const mat4 &projection, const mat4 &view)
{
	camera.projection = projection;
	camera.view = view;
	camera.view_projection = projection * view;
	camera.inv_projection = inverse(projection);
	camera.inv_view = inverse(view);
	camera.inv_view_projection = inverse(camera.view_projection);
	frustum.build_planes(camera.inv_view_projection);

	camera.camera_position = camera.inv_view[3].xyz();
	camera.camera_up = camera.inv_view[1].xyz();
	camera.camera_right = camera.inv_view[0].xyz();
	camera.camera_front = -camera.inv_view[2].xyz();

	// view-space depth of NDC z = 1 (near, reverse-Z) and of z = 0 / 1e-10 (far / "infinite")
	const mat4 &ip = camera.inv_projection;
	auto view_depth_of_ndc_z = [&](float z) {
		float zz = ip[2].z * z + ip[3].z;
		float ww = ip[2].w * z + ip[3].w;
		return -zz / ww;
	};
	bool infinite_z = camera.inv_view_projection[3][3] == 0.0f;
	camera.z_near = view_depth_of_ndc_z(1.0f);
	camera.z_far = view_depth_of_ndc_z(infinite_z ? 1e-10f : 0.0f);
}
} // namespace Granite
