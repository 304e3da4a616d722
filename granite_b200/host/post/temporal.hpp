// temporal.hpp -- TemporalJitter and the TAA resolve pass (renderer/post/temporal.hpp:33-89).
#pragma once

#include <string>
#include <vector>

#include "../math.hpp"
#include "../render_graph.hpp"

namespace Granite
{
class TemporalJitter
{
public:
	enum class Type
	{
		FXAA_2Phase,
		SMAA_T2X,
		TAA_8Phase,
		TAA_16Phase,
		Custom,
		None
	};
	TemporalJitter();
	void init(Type type, vec2 backbuffer_resolution);
	void init_custom(const vec2 *phases, unsigned phase_count, vec2 backbuffer_resolution);
	void step(const mat4 &proj, const mat4 &view);
	const mat4 &get_jitter_matrix() const;
	const mat4 &get_history_view_proj(int frames) const;
	const mat4 &get_history_inv_view_proj(int frames) const;
	const mat4 &get_history_jittered_view_proj(int frames) const;
	const mat4 &get_history_jittered_inv_view_proj(int frames) const;
	const mat4 &get_jittered_projection() const { return saved_jittered_projection; }
	unsigned get_jitter_phase() const { return phase; }
	unsigned get_unmasked_phase() const { return phase; }
	void reset() { phase = 0; }
	Type get_jitter_type() const { return type; }

private:
	unsigned phase = 0;
	unsigned jitter_count = 1;
	std::vector<mat4> jitter_table;
	std::vector<mat4> saved_jittered_view_proj, saved_jittered_inv_view_proj, saved_view_proj, saved_inv_view_proj;
	mat4 saved_jittered_projection;
	Type type = Type::None;
	void init_banks();
	unsigned get_offset_phase(int frames) const;
};

enum class TAAQuality
{
	Low,
	Medium,
	High
};

void setup_taa_resolve(RenderGraph &graph, TemporalJitter &jitter, float scaling_factor, const std::string &input, const std::string &input_depth,
                       const std::string &input_mv, const std::string &output, TAAQuality quality);
} // namespace Granite
