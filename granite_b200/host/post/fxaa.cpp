// fxaa.cpp -- "fxaa" pass builder (renderer/post/fxaa.cpp:28-56).
#include "fxaa.hpp"

namespace Granite
{
void setup_fxaa_postprocess(RenderGraph &graph, const std::string &input, const std::string &output, VkFormat output_format)
{
	// the input is sampled through a UNORM view of its sRGB storage
	graph.get_texture_resource(input).get_attachment_info().flags |= ATTACHMENT_INFO_UNORM_SRGB_ALIAS_BIT;

	auto &fxaa = graph.add_pass("fxaa", RenderGraph::get_default_post_graphics_queue());
	AttachmentInfo fxaa_output;
	fxaa_output.flags |= ATTACHMENT_INFO_SUPPORTS_PREROTATE_BIT;
	fxaa_output.size_class = SizeClass::InputRelative;
	fxaa_output.size_relative_name = input;
	fxaa_output.format = output_format;

	fxaa.add_color_output(output, fxaa_output);
	auto &fxaa_input = fxaa.add_texture_input(input);
	fxaa.set_build_render_pass([&graph, &fxaa, &fxaa_input](Vulkan::CommandBuffer &cmd) {
		auto &input_image = graph.get_physical_texture_resource(fxaa_input);
		auto &output_image = graph.get_physical_texture_resource(*fxaa.get_color_outputs()[0]);
		// cmd.set_unorm_texture(0, 0, input_image): same texels, UNORM interpretation;
		// FXAA_TARGET_SRGB follows the OUTPUT format (fxaa.cpp:50)
		GrbImage in = input_image.as_grb_unorm();
		GrbImage out = output_image.as_grb();
		GrbRows rows = graph.is_sharded() ? graph.get_shard_plan().fxaa : GrbRows{ 0, 0 };
		cmd.check(grb_fxaa(&in, &out, rows, cmd.get_stream_handle()), "grb_fxaa");
	});
}
} // namespace Granite
