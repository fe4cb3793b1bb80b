// scan.cc — FontManager::scan, the `scan` of the recurse command (reference src/commands/recurse.rs:104-133): font
// files, fonts.json manifests (src/commands/recurse.rs:57-63), recursion into directories.
#include "font.h"

#include <algorithm>
#include <cerrno>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <iterator>
#include <sys/stat.h>

namespace vgb {

namespace {
// The subset of JSON a fonts.json needs: an array of objects whose "name" is a string and whose "sources" is an
// array of strings (serde would reject anything else for Vec<FontConfig>, recurse.rs:57-63); unknown keys are skipped.
struct JsonCursor {
	const std::string &s;
	size_t i = 0;
	bool ok = true;
	void ws()
	{
		while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r'))
			++i;
	}
	bool eat(char c)
	{
		ws();
		if (i < s.size() && s[i] == c) {
			++i;
			return true;
		}
		return false;
	}
	static void utf8(std::string &o, uint32_t cp)
	{
		if (cp < 0x80)
			o.push_back((char)cp);
		else if (cp < 0x800) {
			o.push_back((char)(0xC0 | (cp >> 6)));
			o.push_back((char)(0x80 | (cp & 0x3F)));
		} else if (cp < 0x10000) {
			o.push_back((char)(0xE0 | (cp >> 12)));
			o.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
			o.push_back((char)(0x80 | (cp & 0x3F)));
		} else {
			o.push_back((char)(0xF0 | (cp >> 18)));
			o.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
			o.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
			o.push_back((char)(0x80 | (cp & 0x3F)));
		}
	}
	bool hex4(uint32_t &v)
	{
		if (i + 4 > s.size())
			return false;
		v = 0;
		for (int k = 0; k < 4; ++k) {
			const char c = s[i++];
			v <<= 4;
			if (c >= '0' && c <= '9')
				v |= (uint32_t)(c - '0');
			else if (c >= 'a' && c <= 'f')
				v |= (uint32_t)(c - 'a' + 10);
			else if (c >= 'A' && c <= 'F')
				v |= (uint32_t)(c - 'A' + 10);
			else
				return false;
		}
		return true;
	}
	bool string(std::string &out)
	{
		out.clear();
		if (!eat('"'))
			return ok = false;
		while (i < s.size()) {
			const char c = s[i++];
			if (c == '"')
				return true;
			if (c != '\\') {
				out.push_back(c);
				continue;
			}
			if (i >= s.size())
				break;
			const char e = s[i++];
			switch (e) {
			case '"': out.push_back('"'); break;
			case '\\': out.push_back('\\'); break;
			case '/': out.push_back('/'); break;
			case 'b': out.push_back('\b'); break;
			case 'f': out.push_back('\f'); break;
			case 'n': out.push_back('\n'); break;
			case 'r': out.push_back('\r'); break;
			case 't': out.push_back('\t'); break;
			case 'u': {
				uint32_t cp;
				if (!hex4(cp))
					return ok = false;
				if (cp >= 0xD800 && cp < 0xDC00 && i + 6 <= s.size() && s[i] == '\\' && s[i + 1] == 'u') {
					i += 2;
					uint32_t lo;
					if (!hex4(lo) || lo < 0xDC00 || lo > 0xDFFF)
						return ok = false;
					cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
				}
				utf8(out, cp);
				break;
			}
			default: return ok = false;
			}
		}
		return ok = false;
	}
	// skips any value (for keys FontConfig does not have)
	bool skip()
	{
		ws();
		if (i >= s.size())
			return ok = false;
		const char c = s[i];
		if (c == '"') {
			std::string t;
			return string(t);
		}
		if (c == '{' || c == '[') {
			const char close = c == '{' ? '}' : ']';
			++i;
			if (eat(close))
				return true;
			for (;;) {
				if (c == '{') {
					std::string k;
					if (!string(k) || !eat(':'))
						return ok = false;
				}
				if (!skip())
					return false;
				if (eat(','))
					continue;
				return eat(close) ? true : (ok = false);
			}
		}
		const size_t start = i;
		while (i < s.size() && s[i] != ',' && s[i] != '}' && s[i] != ']' && s[i] != ' ' && s[i] != '\n' && s[i] != '\r' && s[i] != '\t')
			++i;
		return i > start ? true : (ok = false);
	}
};

struct FontConfig {
	std::string name;
	std::vector<std::string> sources;
};

bool parse_fonts_json(const std::string &text, std::vector<FontConfig> &out)
{
	JsonCursor c{text};
	if (!c.eat('['))
		return false;
	if (c.eat(']'))
		return true;
	for (;;) {
		if (!c.eat('{'))
			return false;
		FontConfig fc;
		bool has_name = false, has_sources = false;
		if (!c.eat('}')) {
			for (;;) {
				std::string key;
				if (!c.string(key) || !c.eat(':'))
					return false;
				if (key == "name") {
					if (!c.string(fc.name))
						return false;
					has_name = true;
				} else if (key == "sources") {
					if (!c.eat('['))
						return false;
					fc.sources.clear();
					if (!c.eat(']'))
						for (;;) {
							std::string v;
							if (!c.string(v))
								return false;
							fc.sources.push_back(v);
							if (c.eat(','))
								continue;
							if (!c.eat(']'))
								return false;
							break;
						}
					has_sources = true;
				} else if (!c.skip()) {
					return false;
				}
				if (c.eat(','))
					continue;
				if (!c.eat('}'))
					return false;
				break;
			}
		}
		if (!has_name || !has_sources)
			return false; // serde: missing field
		out.push_back(std::move(fc));
		if (c.eat(','))
			continue;
		if (!c.eat(']'))
			return false;
		break;
	}
	c.ws();
	return c.i == text.size();
}

bool has_font_extension(const std::string &path)
{
	const size_t slash = path.find_last_of('/');
	const size_t dot = path.find_last_of('.');
	if (dot == std::string::npos || (slash != std::string::npos && dot < slash))
		return false;
	const std::string ext = path.substr(dot + 1);
	return ext == "ttf" || ext == "otf"; // recurse.rs:106-108 (case-sensitive)
}
} // namespace

bool FontManager::scan(const std::string &path, std::string *err)
{
	struct stat st;
	if (stat(path.c_str(), &st) != 0)
		return true; // neither file nor directory: ignored like the reference's two is_* tests
	if (S_ISREG(st.st_mode)) {
		if (has_font_extension(path))
			return add_path(path, err);
		return true;
	}
	if (!S_ISDIR(st.st_mode))
		return true;
	const std::string manifest = path + "/fonts.json";
	struct stat ms;
	if (stat(manifest.c_str(), &ms) == 0) {
		std::ifstream in(manifest, std::ios::binary);
		std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
		if (!in.good() && !in.eof()) {
			if (err)
				*err = "Failed to read \"" + manifest + "\"";
			return false;
		}
		std::vector<FontConfig> configs;
		if (!parse_fonts_json(text, configs)) {
			if (err)
				*err = "invalid fonts.json: \"" + manifest + "\"";
			return false;
		}
		for (const FontConfig &c : configs) {
			std::vector<std::string> sources;
			for (const std::string &src : c.sources)
				sources.push_back(path + "/" + src);
			if (!add_font_with_name(c.name, sources, err))
				return false;
		}
		return true;
	}
	DIR *d = opendir(path.c_str());
	if (!d) {
		if (err)
			*err = "cannot read directory \"" + path + "\": " + std::strerror(errno);
		return false;
	}
	std::vector<std::string> names;
	while (const dirent *e = readdir(d)) {
		const std::string n = e->d_name;
		if (n != "." && n != "..")
			names.push_back(n);
	}
	closedir(d);
	std::sort(names.begin(), names.end());
	for (const std::string &n : names)
		if (!scan(path + "/" + n, err))
			return false;
	return true;
}

} // namespace vgb
