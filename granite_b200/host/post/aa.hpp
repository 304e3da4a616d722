// aa.hpp -- PostAAType and the two hooks the application calls around the HDR chain
// (renderer/post/aa.hpp:33-63).  FXAA, the three TAA qualities and SMAA 1x (Low .. Ultra) are built; the
// stale FXAA_2Phase and the SMAA_T2X resolve are not.
#pragma once

#include <string>

#include "temporal.hpp"

namespace Granite
{
enum class PostAAType
{
	None,
	FXAA,
	FXAA_2Phase,
	SMAA_Low,
	SMAA_Medium,
	SMAA_High,
	SMAA_Ultra,
	SMAA_Ultra_T2X,
	TAA_Low,
	TAA_Medium,
	TAA_High
};

constexpr bool post_aa_type_is_supported(PostAAType type)
{
	return type == PostAAType::None || type == PostAAType::FXAA || type == PostAAType::TAA_Low || type == PostAAType::TAA_Medium ||
	       type == PostAAType::TAA_High || type == PostAAType::SMAA_Low || type == PostAAType::SMAA_Medium || type == PostAAType::SMAA_High ||
	       type == PostAAType::SMAA_Ultra;
}

// Returns true when a pass was added (the chain input then becomes `output`).
bool setup_before_post_chain_antialiasing(PostAAType type, RenderGraph &graph, TemporalJitter &jitter, float scaling_factor, const std::string &input,
                                          const std::string &input_depth, const std::string &input_mv, const std::string &output);
bool setup_after_post_chain_antialiasing(PostAAType type, RenderGraph &graph, TemporalJitter &jitter, float scaling_factor, const std::string &input,
                                         const std::string &input_depth, const std::string &output);
// FidelityFX FSR 1 from the (scaled-down) post-chain image to the swapchain size (renderer/post/aa.hpp:60,
// aa.cpp:75-174): pass "<output>-scale" (edge-adaptive upscale) and, with use_sharpen, pass "<output>-sharpen"
// (contrast-adaptive sharpening, 0.5 stops as the reference hard-codes).  `output` has the swapchain's size.
bool setup_after_post_chain_upscaling(RenderGraph &graph, const std::string &input, const std::string &output, bool use_sharpen);
} // namespace Granite
