// aa.cpp -- dispatch of PostAAType to the pass builders (renderer/post/aa.cpp:176-290).
#include "aa.hpp"

#include <stdexcept>

#include "fxaa.hpp"
#include "smaa.hpp"

namespace Granite
{
bool setup_before_post_chain_antialiasing(PostAAType type, RenderGraph &graph, TemporalJitter &jitter, float scaling_factor, const std::string &input,
                                          const std::string &input_depth, const std::string &input_mv, const std::string &output)
{
	switch (type)
	{
	case PostAAType::TAA_Low:
		setup_taa_resolve(graph, jitter, scaling_factor, input, input_depth, input_mv, output, TAAQuality::Low);
		return true;
	case PostAAType::TAA_Medium:
		setup_taa_resolve(graph, jitter, scaling_factor, input, input_depth, input_mv, output, TAAQuality::Medium);
		return true;
	case PostAAType::TAA_High:
		setup_taa_resolve(graph, jitter, scaling_factor, input, input_depth, input_mv, output, TAAQuality::High);
		return true;
	default:
		jitter.init(TemporalJitter::Type::None, vec2(0.0f));
		return false;
	}
}

bool setup_after_post_chain_antialiasing(PostAAType type, RenderGraph &graph, TemporalJitter &jitter, float scaling_factor, const std::string &input,
                                         const std::string &input_depth, const std::string &output)
{
	switch (type)
	{
	case PostAAType::FXAA:
		setup_fxaa_postprocess(graph, input, output);
		return true;
	case PostAAType::SMAA_Low:
		setup_smaa_postprocess(graph, jitter, scaling_factor, input, input_depth, output, SMAAPreset::Low);
		return true;
	case PostAAType::SMAA_Medium:
		setup_smaa_postprocess(graph, jitter, scaling_factor, input, input_depth, output, SMAAPreset::Medium);
		return true;
	case PostAAType::SMAA_High:
		setup_smaa_postprocess(graph, jitter, scaling_factor, input, input_depth, output, SMAAPreset::High);
		return true;
	case PostAAType::SMAA_Ultra:
		setup_smaa_postprocess(graph, jitter, scaling_factor, input, input_depth, output, SMAAPreset::Ultra);
		return true;
	case PostAAType::None:
	case PostAAType::TAA_Low:
	case PostAAType::TAA_Medium:
	case PostAAType::TAA_High:
		return false;
	default:
		throw std::logic_error("PostAAType not supported by this executor (FXAA_2Phase and SMAA T2X are not built).");
	}
}
} // namespace Granite
