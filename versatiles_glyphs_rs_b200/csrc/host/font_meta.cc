// font_meta.cc — font naming metadata (mirror of reference src/font/metadata.rs:20-64,84-129,
// src/font/parse_font_name.rs:214-322) and the index files (src/font/index_files.rs:60-139).
// Not on the accelerated path: it only decides output file names and the two JSON indexes.
#include "font.h"

#include <algorithm>
#include <set>
#include <sstream>

namespace vgb {

namespace {

// Script-subset words that are not part of a family name ("Noto Sans Arabic" -> "Noto Sans").
// Data from the reference's SCRIPT_TOKENS table (parse_font_name.rs:1-161), kept as one string.
const char *const kScriptTokens =
	    "aboriginal adlam albanian anatolian arabic aramaic armenian avestan balinese bamum bassa batak bengali "
	    "bhaiksuki brahmi buginese buhid canadian carian caucasian chakma cham cherokee chiki cin coptic "
	    "cuneiform cypriot deseret devanagari duployan egyptian elbasan elymaic ethiopic georgian glagolitic "
	    "gondi gothic grantha gujarati gunjala gurmukhi hanifi hanunoo hatran hau hebrew hieroglyphs hmong "
	    "hungarian imperial indic inscriptional italic javanese jp kaithi kannada kayah kharoshthi khmer khojki "
	    "khudawadi kikakui kr lao le lepcha li limbu linear lisu lue lycian lydian mahajani malayalam mandaic "
	    "manichaean marchen masaram mayan mayek medefaidrin meetei mende meroitic miao modi mongolian mro multani "
	    "myanmar nabataean new newa nko north numbers nushu ogham ol old oriya osage osmanya pa pahawh pahlavi "
	    "palmyrene parthian pau permic persian phags phoenician psalter rejang rohingya runic samaritan "
	    "saurashtra sc sharada shavian siddham sinhala sogdian sompeng sora south soyombo square sundanese syloti "
	    "symbols syriac tagalog tagbanwa tai takri tamil tangut tc telugu thaana thai tibetan tifinagh tirhuta "
	    "turkic ugaritic vah vai wancho warang yi zanabazar ";

bool is_script_token(const std::string &t)
{
	static const std::set<std::string> tokens = [] {
		std::set<std::string> s;
		std::istringstream in(kScriptTokens);
		for (std::string w; in >> w;)
			s.insert(w);
		return s;
	}();
	return tokens.count(t) != 0;
}

std::string lower(std::string s)
{
	for (char &c : s)
		if (c >= 'A' && c <= 'Z')
			c = (char)(c - 'A' + 'a');
	return s;
}

bool has(const std::string &s, const char *needle) { return s.find(needle) != std::string::npos; }

// parse_font_name.rs:295-322 — most specific keyword first
uint16_t find_weight(const std::string &s)
{
	if (has(s, "hairline") || has(s, "thin"))
		return 100;
	if (has(s, "extralight") || has(s, "ultralight"))
		return 200;
	if (has(s, "light"))
		return 300;
	if (has(s, "regular") || has(s, "normal") || has(s, "book"))
		return 400;
	if (has(s, "medium"))
		return 500;
	if (has(s, "demibold") || has(s, "semibold"))
		return 600;
	if (has(s, "bold"))
		return (has(s, "extra") || has(s, "ultra")) ? 800 : 700;
	if (has(s, "black") || has(s, "heavy"))
		return 900;
	return 400;
}

} // namespace

// serde_json's string escaping: \" \\ \b \f \n \r \t, other control characters as \u00XX
std::string json_escape(const std::string &s)
{
	std::string o;
	for (unsigned char c : s) {
		switch (c) {
		case '"': o += "\\\""; break;
		case '\\': o += "\\\\"; break;
		case '\b': o += "\\b"; break;
		case '\f': o += "\\f"; break;
		case '\n': o += "\\n"; break;
		case '\r': o += "\\r"; break;
		case '\t': o += "\\t"; break;
		default:
			if (c < 0x20) {
				char buf[8];
				std::snprintf(buf, sizeof(buf), "\\u%04x", c);
				o += buf;
			} else {
				o.push_back((char)c);
			}
		}
	}
	return o;
}

// parse_font_name.rs:214-293
void parse_font_name(const std::string &family, const std::string &ps_name, std::string &out_family, std::string &style,
                     uint16_t &weight, std::string &width)
{
	style = "normal";
	weight = 400;
	width = "normal";
	const size_t dash = ps_name.rfind('-');
	const std::string suffix = lower(dash == std::string::npos ? ps_name : ps_name.substr(dash + 1));
	if (has(suffix, "italic"))
		style = "italic";
	const uint16_t ps_weight = find_weight(suffix);
	if (ps_weight != 400)
		weight = ps_weight;

	std::vector<std::string> tokens;
	{
		std::istringstream in(family); // split_whitespace
		for (std::string w; in >> w;)
			tokens.push_back(w);
	}
	std::string fam;
	for (size_t i = 0; i < tokens.size(); ++i) {
		const std::string t = lower(tokens[i]);
		if (i + 1 < tokens.size() && t == "extra" && lower(tokens[i + 1]) == "condensed") {
			width = "extra-condensed";
			++i;
			continue;
		}
		if (t == "semicondensed" || t == "semi-condensed") {
			width = "semi-condensed";
			continue;
		}
		if (t == "condensed") {
			width = "condensed";
			continue;
		}
		if (is_script_token(t))
			continue;
		const uint16_t w = find_weight(t);
		if (w != 400) {
			if (ps_weight == 400)
				weight = w;
			continue;
		}
		if (!fam.empty())
			fam.push_back(' ');
		fam += tokens[i];
	}
	out_family = fam;
}

// metadata.rs:29-55
std::string FontMetadata::generate_name() const
{
	std::string n = family;
	if (width != "normal")
		n += " " + width;
	const char *w = "Unknown";
	switch (weight) {
	case 100: w = "Thin"; break;
	case 200: w = "ExtraLight"; break;
	case 300: w = "Light"; break;
	case 400: w = "Regular"; break;
	case 500: w = "Medium"; break;
	case 600: w = "SemiBold"; break;
	case 700: w = "Bold"; break;
	case 800: w = "ExtraBold"; break;
	case 900: w = "Black"; break;
	default: break;
	}
	n += std::string(" ") + w;
	if (style != "normal")
		n += " " + style;
	return n;
}

// metadata.rs:84-129 (names: family = id 1, PostScript name = id 6)
FontMetadata FontMetadata::from_face(const Face &face)
{
	FontMetadata m;
	m.name = face.name(1);
	parse_font_name(m.name, face.name(6), m.family, m.style, m.weight, m.width);
	m.codepoints = face.codepoints();
	return m;
}

// index_files.rs:60-95 — 16-code-point blocks, consecutive blocks merged, upper-case hex
std::string encode_codeblocks(const std::vector<uint32_t> &codepoints)
{
	std::vector<uint32_t> blocks;
	for (uint32_t cp : codepoints)
		blocks.push_back(cp >> 4);
	std::sort(blocks.begin(), blocks.end());
	blocks.erase(std::unique(blocks.begin(), blocks.end()), blocks.end());
	std::string out;
	char buf[32];
	for (size_t i = 0; i < blocks.size();) {
		size_t j = i;
		while (j + 1 < blocks.size() && blocks[j + 1] == blocks[j] + 1)
			++j;
		if (!out.empty())
			out.push_back(',');
		if (i == j)
			std::snprintf(buf, sizeof(buf), "%X", blocks[i]);
		else
			std::snprintf(buf, sizeof(buf), "%X-%X", blocks[i], blocks[j]);
		out += buf;
		i = j + 1;
	}
	return out;
}

// index_files.rs:115-139 — serde_json::to_vec_pretty of [{name, faces:[{id,style,weight,width,codeblocks}]}]
std::string build_font_families_json(const std::map<std::string, FontWrapper> &fonts, std::string *err)
{
	struct Face_ {
		std::string id, style, width, codeblocks;
		uint16_t weight;
	};
	std::map<std::string, std::vector<Face_>> families; // sorted by family name (index_files.rs:136)
	for (const auto &kv : fonts) {
		if (kv.second.files().empty()) {
			if (err)
				*err = "FontWrapper has no files"; // wrapper.rs:80-85
			return std::string();
		}
		const FontMetadata &m = kv.second.files().front()->metadata;
		families[m.family].push_back(Face_{kv.first, m.style, m.width, encode_codeblocks(m.codepoints), m.weight});
	}
	if (families.empty())
		return "[]";
	std::string s = "[\n";
	size_t fi = 0;
	for (const auto &fam : families) {
		s += "  {\n    \"name\": \"" + json_escape(fam.first) + "\",\n    \"faces\": [";
		if (fam.second.empty()) {
			s += "]\n";
		} else {
			s += "\n";
			for (size_t k = 0; k < fam.second.size(); ++k) {
				const Face_ &f = fam.second[k];
				s += "      {\n        \"id\": \"" + json_escape(f.id) + "\",\n        \"style\": \"" + json_escape(f.style) +
				     "\",\n        \"weight\": " + std::to_string(f.weight) + ",\n        \"width\": \"" + json_escape(f.width) +
				     "\",\n        \"codeblocks\": \"" + f.codeblocks + "\"\n      }";
				s += (k + 1 < fam.second.size()) ? ",\n" : "\n";
			}
			s += "    ]\n";
		}
		s += (++fi < families.size()) ? "  },\n" : "  }\n";
	}
	s += "]";
	return s;
}

} // namespace vgb
