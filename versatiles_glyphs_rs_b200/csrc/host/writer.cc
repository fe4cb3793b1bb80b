// writer.cc — Writer (mirror of reference src/writer/mod.rs:21-96): directory sink (writer/file.rs), in-memory
// recorder (writer/dummy.rs + contents), ustar stream (writer/tar.rs:49-137).
#include "font.h"

#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sys/stat.h>

namespace vgb {

// ---- Writer ------------------------------------------------------------------------------------------
Writer Writer::new_file(const std::string &folder)
{
	Writer w;
	w.to_disk_ = true;
	w.folder_ = folder;
	return w;
}
Writer Writer::new_memory() { return Writer(); }

Writer Writer::new_tar_memory()
{
	Writer w;
	w.to_tar_ = true;
	return w;
}

Writer Writer::new_tar(const std::string &path)
{
	Writer w;
	w.to_tar_ = true;
	w.folder_ = path;
	std::FILE *f = std::fopen(path.c_str(), "wb");
	if (f) {
		std::shared_ptr<bool> closed(new bool(false));
		w.tar_closed_ = closed;
		w.tar_file_ = std::shared_ptr<std::FILE>(f, [closed](std::FILE *p) {
			if (!*closed)
				std::fclose(p); // (a writer dropped without finish(): best effort, writer/mod.rs:83-96)
		});
	}
	return w;
}

bool Writer::tar_put(const uint8_t *p, size_t n, std::string *err)
{
	if (!folder_.empty()) {
		if (!tar_file_ || (tar_closed_ && *tar_closed_) || (n && std::fwrite(p, 1, n, tar_file_.get()) != n)) {
			if (err)
				*err = "writing tar \"" + folder_ + "\" failed";
			return false;
		}
		return true;
	}
	tar_.insert(tar_.end(), p, p + n);
	return true;
}

// writer/tar.rs:49-99 — one 512-byte ustar header; octal fields are zero-filled and end with a space
bool Writer::tar_header(const std::string &path, uint64_t size, uint64_t mode, char typeflag, std::string *err)
{
	uint8_t h[512];
	std::memset(h, 0, sizeof(h));
	if (path.size() > 100) { // tar.rs:160-172
		if (err)
			*err = "tar header field overflow: \"" + path + "\" is " + std::to_string(path.size()) + " bytes, max 100";
		return false;
	}
	std::memcpy(h, path.data(), path.size());
	auto octal = [&](size_t off, size_t len, uint64_t v) { // tar.rs:147-156
		h[off + len - 1] = ' ';
		for (size_t i = len - 1; i-- > 0;) {
			h[off + i] = (uint8_t)('0' + (v & 7));
			v >>= 3;
		}
	};
	octal(100, 8, mode);
	octal(108, 8, 0);
	octal(116, 8, 0);
	octal(124, 12, size);
	octal(136, 12, (uint64_t)std::chrono::duration_cast<std::chrono::seconds>(std::chrono::system_clock::now().time_since_epoch()).count());
	h[156] = (uint8_t)typeflag;
	std::memcpy(h + 257, "ustar\0", 6);
	std::memcpy(h + 263, "00", 2);
	std::memset(h + 148, ' ', 8);
	uint32_t sum = 0;
	for (uint8_t b : h)
		sum += b;
	octal(148, 8, sum);
	return tar_put(h, sizeof(h), err);
}

static bool mkdirs(const std::string &path, std::string *err)
{
	std::string cur;
	for (size_t i = 0; i <= path.size(); ++i) {
		if (i == path.size() || path[i] == '/') {
			if (!cur.empty() && ::mkdir(cur.c_str(), 0755) != 0 && errno != EEXIST) {
				if (err)
					*err = "mkdir " + cur + ": " + std::strerror(errno);
				return false;
			}
		}
		if (i < path.size())
			cur.push_back(path[i]);
	}
	return true;
}

bool Writer::write_file(const std::string &filename, const uint8_t *bytes, size_t len, std::string *err)
{
	__atomic_fetch_add(&bytes_written_, (uint64_t)len, __ATOMIC_RELAXED);
	if (to_tar_) { // writer/tar.rs:101-120
		static const uint8_t zeros[512] = {0};
		if (!tar_header(filename, len, 0644, '0', err) || !tar_put(bytes, len, err))
			return false;
		const size_t rem = len % 512;
		return rem == 0 || tar_put(zeros, 512 - rem, err);
	}
	if (!to_disk_) {
		Entry e;
		e.name = filename;
		e.bytes.assign(bytes, bytes + len);
		entries_.push_back(std::move(e));
		return true;
	}
	const std::string path = folder_ + "/" + filename;
	std::FILE *f = std::fopen(path.c_str(), "wb");
	if (!f) {
		if (err)
			*err = "open " + path + ": " + std::strerror(errno);
		return false;
	}
	bool ok = len == 0 || std::fwrite(bytes, 1, len, f) == len;
	// (buffered data reaches the file at close: a full disk or an I/O error may only show up here — std::fs::write in the
	// reference, writer/file.rs:42, reports it)
	if (std::fclose(f) != 0)
		ok = false;
	if (!ok && err)
		*err = "write " + path + " failed: " + std::strerror(errno);
	return ok;
}

bool Writer::write_file(const std::string &filename, std::vector<uint8_t> &&bytes, std::string *err)
{
	if (to_disk_ || to_tar_)
		return write_file(filename, bytes.data(), bytes.size(), err);
	__atomic_fetch_add(&bytes_written_, (uint64_t)bytes.size(), __ATOMIC_RELAXED);
	Entry e;
	e.name = filename;
	e.bytes = std::move(bytes);
	entries_.push_back(std::move(e));
	return true;
}

bool Writer::write_directory(const std::string &dirname, std::string *err)
{
	if (to_tar_) { // writer/tar.rs:122-126
		if (dirname.empty() || dirname.back() != '/') {
			if (err)
				*err = "dirname must end with a slash";
			return false;
		}
		return tar_header(dirname, 0, 0755, '5', err);
	}
	if (!to_disk_) {
		Entry e;
		e.name = dirname;
		e.is_dir = true;
		entries_.push_back(std::move(e));
		return true;
	}
	return mkdirs(folder_ + "/" + dirname, err);
}

bool Writer::finish(std::string *err)
{
	if (finished_) // writer/mod.rs:67-73
		return true;
	finished_ = true;
	if (to_tar_) { // writer/tar.rs:133-137: two zero blocks
		static const uint8_t zeros[1024] = {0};
		if (!tar_put(zeros, sizeof(zeros), err))
			return false;
		if (tar_file_ && tar_closed_ && !*tar_closed_) {
			// BufWriter::flush + close (writer/tar.rs:133-137): errors that only surface now are errors of finish()
			std::FILE *f = tar_file_.get();
			const bool flushed = std::fflush(f) == 0;
			const bool closed = std::fclose(f) == 0;
			*tar_closed_ = true;
			if (!flushed || !closed) {
				if (err)
					*err = std::string("closing the tar file failed: ") + std::strerror(errno);
				return false;
			}
		}
	}
	return true;
}

} // namespace vgb
