// taa_ini.cu — settings-file compatibility: taa<CF>::writeSettingsToIni / readSettingsFromIni (source/taa.hpp:1198-1339)
// over the INI text the reference's mINI dependency (external/include/mini/ini.h, not case sensitive) reads and generates.
// Pure host code: the same section and key names, the same value formats (source/IniUtil.cpp:52-56, 104-109), the same
// "absent or empty value keeps the current setting" rule, the same key order when writing.
//
// The format, as mINI parses it (ini.h:275-322): lines are trimmed; a line starting with ';' is a comment; "[name]" opens a section
// (anything after a ';' on that line is dropped); otherwise the first '=' not written as "\=" splits key and value, both trimmed;
// section and key names are compared in lower case; later duplicates overwrite earlier ones. generate() writes "[section]" and then its
// "key=value" lines with the lower-cased names, one per line, no blank lines, no newline after the last line
// (tests/golden/taa_settings_written.ini is what the reference's own code produces).
#include <cctype>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "taa_ctx.h"

namespace {

std::string trim(const std::string& s) {
	size_t a = 0, b = s.size();
	while (a < b && std::isspace((unsigned char)s[a])) ++a;
	while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
	return s.substr(a, b - a);
}
std::string lower(std::string s) {
	for (char& c : s) c = (char)std::tolower((unsigned char)c);
	return s;
}

using Section = std::map<std::string, std::string>;
using Ini = std::map<std::string, Section>;

Ini parse(const char* text) {
	Ini ini;
	std::string section;
	const char* p = text;
	while (*p) {
		const char* e = p;
		while (*e && *e != '\n') ++e;
		std::string line = trim(std::string(p, e));
		p = *e ? e + 1 : e;
		if (line.empty() || line[0] == ';') continue;
		if (line[0] == '[') {
			const size_t c = line.find_first_of(';');
			if (c != std::string::npos) line = line.substr(0, c);
			const size_t r = line.find_last_of(']');
			if (r != std::string::npos) {
				section = lower(trim(line.substr(1, r - 1)));
				ini[section];
				continue;
			}
		}
		std::string norm = line;  // "\=" is an escaped '=' inside a key
		for (size_t i = 0; i + 1 < norm.size(); ++i)
			if (norm[i] == '\\' && norm[i + 1] == '=') norm[i] = norm[i + 1] = ' ';
		const size_t eq = norm.find_first_of('=');
		if (eq == std::string::npos) continue;  // PDATA_UNKNOWN: ignored
		std::string key = trim(line.substr(0, eq));
		for (size_t i = key.find("\\="); i != std::string::npos; i = key.find("\\=", i + 1)) key.replace(i, 2, "=");
		ini[section][lower(key)] = trim(line.substr(eq + 1));
	}
	return ini;
}

struct Reader {
	const Ini& ini;
	std::string sec;
	std::string err;
	const std::string* find(const std::string& name) const {
		auto s = ini.find(lower(sec));
		if (s == ini.end()) return nullptr;
		auto k = s->second.find(lower(name));
		if (k == s->second.end() || k->second.empty()) return nullptr;  // `if (ini[section][name] != "")`, IniUtil.cpp:104-109
		return &k->second;
	}
	void bad(const std::string& name, const std::string& v) {
		if (err.empty()) err = "[" + sec + "] " + name + " = '" + v + "': not a number";
	}
	void rBool(const std::string& name, taa_bool32& val) {  // iniReadBool: "1" or "true"
		if (const std::string* v = find(name)) val = (*v == "1" || *v == "true") ? 1u : 0u;
	}
	void rBool32(const std::string& name, taa_bool32& val) {  // iniReadBool32: std::stoul
		if (const std::string* v = find(name)) {
			char* end = nullptr;
			const unsigned long x = std::strtoul(v->c_str(), &end, 10);
			if (end == v->c_str()) bad(name, *v); else val = (taa_bool32)x;
		}
	}
	void rInt(const std::string& name, int32_t& val) {  // iniReadInt: std::stol
		if (const std::string* v = find(name)) {
			char* end = nullptr;
			const long x = std::strtol(v->c_str(), &end, 10);
			if (end == v->c_str()) bad(name, *v); else val = (int32_t)x;
		}
	}
	void rFloat(const std::string& name, float& val) {  // iniReadFloat: std::stof
		if (const std::string* v = find(name)) {
			char* end = nullptr;
			const float x = std::strtof(v->c_str(), &end);
			if (end == v->c_str()) bad(name, *v); else val = x;
		}
	}
	void rVec4(const std::string& name, float* v) { rFloat(name + ".x", v[0]); rFloat(name + ".y", v[1]); rFloat(name + ".z", v[2]); rFloat(name + ".w", v[3]); }
	void rIVec4(const std::string& name, int32_t* v) { rInt(name + ".x", v[0]); rInt(name + ".y", v[1]); rInt(name + ".z", v[2]); rInt(name + ".w", v[3]); }
};

struct Writer {  // INIGenerator (ini.h): one line per section header / pair, '\n' between lines, none after the last
	std::string out;
	void line(const std::string& l) {
		if (!out.empty()) out += "\n";
		out += l;
	}
	void section(const std::string& name) { line("[" + lower(name) + "]"); }
	void kv(const std::string& name, const std::string& v) { line(lower(name) + "=" + v); }
	void wBool(const std::string& n, taa_bool32 v) { kv(n, v ? "1" : "0"); }                         // iniWriteBool
	void wBool32(const std::string& n, taa_bool32 v) { kv(n, std::to_string((unsigned int)v)); }     // iniWriteBool32
	void wInt(const std::string& n, int32_t v) { kv(n, std::to_string(v)); }                         // iniWriteInt
	void wFloat(const std::string& n, float v) { kv(n, std::to_string(v)); }                         // iniWriteFloat: "%f"
	void wVec4(const std::string& n, const float* v) { wFloat(n + ".x", v[0]); wFloat(n + ".y", v[1]); wFloat(n + ".z", v[2]); wFloat(n + ".w", v[3]); }
	void wIVec4(const std::string& n, const int32_t* v) { wInt(n + ".x", v[0]); wInt(n + ".y", v[1]); wInt(n + ".z", v[2]); wInt(n + ".w", v[3]); }
};

thread_local std::string g_ini_error;

}  // namespace

extern "C" {

// writeSettingsToIni, taa.hpp:1198-1265. Returns the number of bytes of the text including the terminating NUL; writes at most `cap`.
TAA_API int32_t taa_settings_write_ini(const TaaParameters params[2], const taa_invokee_settings* s, const TaaPostProcessPush* pp, char* out, int32_t cap) {
	if (!params || !s || !pp) return TAA_E_INVALID_ARG;
	Writer w;
	for (int i = 0; i < 2; ++i) {
		const TaaParameters& p = params[i];
		w.section("TAA_Param_" + std::to_string(i));
		w.wBool32("mPassThrough", p.mPassThrough);
		w.wInt("mColorClampingOrClipping", p.mColorClampingOrClipping);
		w.wBool32("mShapedNeighbourhood", p.mShapedNeighbourhood);
		w.wBool32("mVarianceClipping", p.mVarianceClipping);
		w.wFloat("mVarClipGamma", p.mVarClipGamma);
		w.wBool32("mUseYCoCg", p.mUseYCoCg);
		w.wBool32("mLumaWeightingLottes", p.mLumaWeightingLottes);
		w.wBool32("mDepthCulling", p.mDepthCulling);
		w.wBool32("mRejectOutside", p.mRejectOutside);
		w.wBool32("mUnjitterNeighbourhood", p.mUnjitterNeighbourhood);
		w.wBool32("mUnjitterCurrentSample", p.mUnjitterCurrentSample);
		w.wFloat("mUnjitterFactor", p.mUnjitterFactor);
		w.wFloat("mAlpha", p.mAlpha);
		w.wFloat("mMinAlpha", p.mMinAlpha);
		w.wFloat("mMaxAlpha", p.mMaxAlpha);
		w.wFloat("mRejectionAlpha", p.mRejectionAlpha);
		w.wInt("mUseVelocityVectors", p.mUseVelocityVectors);
		w.wInt("mVelocitySampleMode", p.mVelocitySampleMode);
		w.wInt("mInterpolationMode", p.mInterpolationMode);
		w.wBool32("mToneMapLumaKaris", p.mToneMapLumaKaris);
		w.wBool32("mAddNoise", p.mAddNoise);
		w.wFloat("mNoiseFactor", p.mNoiseFactor);
		w.wBool32("mReduceBlendNearClamp", p.mReduceBlendNearClamp);
		w.wBool32("mDynamicAntiGhosting", p.mDynamicAntiGhosting);
		w.wVec4("mDebugMask", p.mDebugMask);
		w.wInt("mDebugMode", p.mDebugMode);
		w.wFloat("mDebugScale", p.mDebugScale);
		w.wBool32("mDebugCenter", p.mDebugCenter);
		w.wBool32("mDebugToScreenOutput", p.mDebugToScreenOutput);
		w.wBool32("mVelBasedAlpha", p.mVelBasedAlpha);
		w.wFloat("mVelBasedAlphaMax", p.mVelBasedAlphaMax);
		w.wFloat("mVelBasedAlphaFactor", p.mVelBasedAlphaFactor);
		w.wBool32("mRayTraceAugment", p.mRayTraceAugment);
	}
	w.section("TAA_Primary");
	w.wBool("mTaaEnabled", s->mTaaEnabled);
	w.wInt("mSampleDistribution", s->jitter.mSampleDistribution);
	w.wInt("mSharpener", s->mSharpener);
	w.wFloat("mSharpenFactor", s->mSharpenFactor);
	w.wBool("mSplitScreen", s->mSplitScreen);
	w.wInt("mSplitX", s->mSplitX);
	w.wInt("mFixedJitterIndex", s->jitter.mFixedJitterIndex);
	w.wFloat("mJitterExtraScale", s->jitter.mJitterExtraScale);
	w.wInt("mJitterSlowMotion", s->jitter.mJitterSlowMotion);
	w.wFloat("mJitterRotateDegrees", s->jitter.mJitterRotateDegrees);
	w.wBool("mResetHistoryOnChange", s->mResetHistoryOnChange);
	const int n = s->jitter.mDebugSampleOffsets ? s->jitter.mDebugSampleOffsetsCount : 0;
	w.wInt("mDebugSampleOffsets.size", n);
	for (int i = 0; i < n; ++i) {
		w.wFloat("mDebugSampleOffsets_" + std::to_string(i) + ".x", s->jitter.mDebugSampleOffsets[2 * i]);
		w.wFloat("mDebugSampleOffsets_" + std::to_string(i) + ".y", s->jitter.mDebugSampleOffsets[2 * i + 1]);
	}
	w.section("TAA_Postprocess");
	w.wBool("mPostProcessEnabled", s->mPostProcessEnabled);
	w.wBool32("zoom", pp->zoom);
	w.wBool32("showZoomBox", pp->showZoomBox);
	w.wIVec4("zoomSrcLTWH", pp->zoomSrcLTWH);
	w.wIVec4("zoomDstLTWH", pp->zoomDstLTWH);
	const int32_t need = (int32_t)w.out.size() + 1;
	if (out && cap > 0) {
		const int32_t n_copy = need <= cap ? need - 1 : cap - 1;
		memcpy(out, w.out.data(), (size_t)n_copy);
		out[n_copy] = '\0';
	}
	return need;
}

// readSettingsFromIni, taa.hpp:1267-1339. `offsets` receives mDebugSampleOffsets (the current ones, resized to max(1, size) vec2s as the
// reference's vector is, then overwritten key by key) and is what s->jitter.mDebugSampleOffsets points to afterwards; offsets_cap counts vec2s.
TAA_API int taa_settings_read_ini(const char* text, TaaParameters params[2], taa_invokee_settings* s, TaaPostProcessPush* pp, float* offsets,
                                  int32_t offsets_cap) {
	g_ini_error.clear();
	if (!text || !params || !s || !pp) { g_ini_error = "null argument"; return TAA_E_INVALID_ARG; }
	const Ini ini = parse(text);
	Reader r{ini, "", ""};
	for (int i = 0; i < 2; ++i) {
		TaaParameters& p = params[i];
		r.sec = "TAA_Param_" + std::to_string(i);
		r.rBool32("mPassThrough", p.mPassThrough);
		r.rInt("mColorClampingOrClipping", p.mColorClampingOrClipping);
		r.rBool32("mShapedNeighbourhood", p.mShapedNeighbourhood);
		r.rBool32("mVarianceClipping", p.mVarianceClipping);
		r.rFloat("mVarClipGamma", p.mVarClipGamma);
		r.rBool32("mUseYCoCg", p.mUseYCoCg);
		r.rBool32("mLumaWeightingLottes", p.mLumaWeightingLottes);
		r.rBool32("mDepthCulling", p.mDepthCulling);
		r.rBool32("mRejectOutside", p.mRejectOutside);
		r.rBool32("mUnjitterNeighbourhood", p.mUnjitterNeighbourhood);
		r.rBool32("mUnjitterCurrentSample", p.mUnjitterCurrentSample);
		r.rFloat("mUnjitterFactor", p.mUnjitterFactor);
		r.rFloat("mAlpha", p.mAlpha);
		r.rFloat("mMinAlpha", p.mMinAlpha);
		r.rFloat("mMaxAlpha", p.mMaxAlpha);
		r.rFloat("mRejectionAlpha", p.mRejectionAlpha);
		r.rInt("mUseVelocityVectors", p.mUseVelocityVectors);
		r.rInt("mVelocitySampleMode", p.mVelocitySampleMode);
		r.rInt("mInterpolationMode", p.mInterpolationMode);
		r.rBool32("mToneMapLumaKaris", p.mToneMapLumaKaris);
		r.rBool32("mAddNoise", p.mAddNoise);
		r.rFloat("mNoiseFactor", p.mNoiseFactor);
		r.rBool32("mReduceBlendNearClamp", p.mReduceBlendNearClamp);
		r.rBool32("mDynamicAntiGhosting", p.mDynamicAntiGhosting);
		r.rBool32("mVelBasedAlpha", p.mVelBasedAlpha);
		r.rFloat("mVelBasedAlphaMax", p.mVelBasedAlphaMax);
		r.rFloat("mVelBasedAlphaFactor", p.mVelBasedAlphaFactor);
		r.rVec4("mDebugMask", p.mDebugMask);
		r.rInt("mDebugMode", p.mDebugMode);
		r.rFloat("mDebugScale", p.mDebugScale);
		r.rBool32("mDebugCenter", p.mDebugCenter);
		r.rBool32("mDebugToScreenOutput", p.mDebugToScreenOutput);
		r.rBool32("mRayTraceAugment", p.mRayTraceAugment);
	}
	r.sec = "TAA_Primary";
	r.rBool("mTaaEnabled", s->mTaaEnabled);
	r.rInt("mSampleDistribution", s->jitter.mSampleDistribution);
	r.rInt("mSharpener", s->mSharpener);
	r.rFloat("mSharpenFactor", s->mSharpenFactor);
	r.rBool("mSplitScreen", s->mSplitScreen);
	r.rInt("mSplitX", s->mSplitX);
	r.rInt("mFixedJitterIndex", s->jitter.mFixedJitterIndex);
	r.rFloat("mJitterExtraScale", s->jitter.mJitterExtraScale);
	r.rInt("mJitterSlowMotion", s->jitter.mJitterSlowMotion);
	r.rFloat("mJitterRotateDegrees", s->jitter.mJitterRotateDegrees);
	r.rBool("mResetHistoryOnChange", s->mResetHistoryOnChange);
	int32_t n_samples = 0;
	r.rInt("mDebugSampleOffsets.size", n_samples);
	if (offsets) {
		const int32_t n = n_samples > 1 ? n_samples : 1;  // mDebugSampleOffsets.resize(std::max(1, nSamples), glm::vec2(0)):
		if (n > offsets_cap) {                            // the entries there are stay, new ones are zero
			g_ini_error = "mDebugSampleOffsets.size = " + std::to_string(n_samples) + " exceeds the caller's buffer (" + std::to_string(offsets_cap) + " vec2)";
			return TAA_E_INVALID_ARG;
		}
		const int32_t n_old = s->jitter.mDebugSampleOffsets ? s->jitter.mDebugSampleOffsetsCount : 0;
		if (s->jitter.mDebugSampleOffsets != offsets)
			for (int32_t i = 0; i < 2 * n && i < 2 * n_old; ++i) offsets[i] = s->jitter.mDebugSampleOffsets[i];
		for (int32_t i = 2 * (n_old < n ? n_old : n); i < 2 * n; ++i) offsets[i] = 0.f;
		for (int32_t i = 0; i < n_samples; ++i) {
			r.rFloat("mDebugSampleOffsets_" + std::to_string(i) + ".x", offsets[2 * i]);
			r.rFloat("mDebugSampleOffsets_" + std::to_string(i) + ".y", offsets[2 * i + 1]);
		}
		s->jitter.mDebugSampleOffsets = offsets;
		s->jitter.mDebugSampleOffsetsCount = n;
	}
	r.sec = "TAA_Postprocess";
	r.rBool("mPostProcessEnabled", s->mPostProcessEnabled);
	r.rBool32("zoom", pp->zoom);
	r.rBool32("showZoomBox", pp->showZoomBox);
	r.rIVec4("zoomSrcLTWH", pp->zoomSrcLTWH);
	r.rIVec4("zoomDstLTWH", pp->zoomDstLTWH);
	if (!r.err.empty()) {  // std::stoul / stol / stof would have thrown std::invalid_argument
		g_ini_error = r.err;
		return TAA_E_INVALID_ARG;
	}
	return TAA_OK;
}

TAA_API const char* taa_settings_ini_last_error(void) { return g_ini_error.c_str(); }

}  // extern "C"
