// ingest.cu -- FASTA / FASTQ ingest in front of the batched C ABI (host code only).
//
// Replaces the reference's driver loop  "read a FASTA record, insertSeq(filter, seq, h, k)"
// (swig/writeBloom_rolling.cpp:19-59, BloomFilterUtil.h:10-17): the file is cut into regions, one parser
// thread per region turns its lines into flat batches (concatenated bases + offsets) in pinned buffers,
// and one thread streams the batches through btlbf_insert_seqs_async / btlbf_contains_seqs_async, so that
// parsing, H2D copies and kernels overlap.  A sequence that does not fit a batch (or that crosses a region
// boundary) continues in the next batch as a new piece that starts with the previous k-1 bases, so every
// window is visited exactly once.  Bytes that are not bases simply make their windows invalid on the
// device, exactly as in the reference's iterator.
#include "../../include/btlbf.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fcntl.h>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

extern "C" int btlbf_set_error(int code, const char* msg); // capi.cu: stores the thread-local error text

namespace {

int failf(int code, const char* fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	return btlbf_set_error(code, buf);
}

bool pread_all(int fd, void* dst, size_t n, uint64_t pos)
{
	char* p = (char*)dst;
	while (n) {
		ssize_t r = pread(fd, p, n, (off_t)pos);
		if (r <= 0)
			return false;
		p += r;
		n -= (size_t)r;
		pos += (uint64_t)r;
	}
	return true;
}

enum LineState { LINE_START, IN_HEADER, IN_SEQ, IN_SKIP };

} // namespace

// One reader = one region [lo, hi) of the file: it owns the lines (FASTA) / records (FASTQ) that start in it.
struct btlbf_seqfile
{
	int fd = -1;
	bool fastq = false;
	unsigned overlap = 0;
	uint64_t size = 0, pos = 0, hi = 0;
	std::vector<char> raw;
	size_t raw_len = 0, raw_off = 0;
	LineState st = LINE_START;
	int fq_line = 0;        // FASTQ: 0 header, 1 sequence, 2 '+', 3 quality
	bool seq_open = false;  // the current record's sequence may continue
	bool piece_open = false;
	std::string tail;       // last <= overlap bases of the open sequence (start of its next piece)
	bool done = false;
	std::string path;

	bool fill()
	{
		if (raw_off < raw_len)
			return true;
		if (pos >= size)
			return false;
		size_t want = (size_t)std::min<uint64_t>(size - pos, (uint64_t)8 << 20);
		raw.resize(want);
		if (!pread_all(fd, raw.data(), want, pos))
			return false;
		raw_len = want;
		raw_off = 0;
		return true;
	}
	uint64_t file_pos() const { return pos + raw_off; }
	void advance_block()
	{
		pos += raw_len;
		raw_len = raw_off = 0;
	}
};

namespace {

// first line start >= p (p itself when it follows a newline or is 0)
bool next_line_start(btlbf_seqfile* r, uint64_t p, uint64_t* out)
{
	if (p == 0) {
		*out = 0;
		return true;
	}
	char buf[65536];
	uint64_t q = p - 1;
	while (q < r->size) {
		size_t n = (size_t)std::min<uint64_t>(sizeof buf, r->size - q);
		if (!pread_all(r->fd, buf, n, q))
			return false;
		const void* nl = memchr(buf, '\n', n);
		if (nl) {
			*out = q + (uint64_t)((const char*)nl - buf) + 1;
			return true;
		}
		q += n;
	}
	*out = r->size;
	return true;
}

char byte_at(btlbf_seqfile* r, uint64_t p)
{
	char c = 0;
	if (p < r->size)
		pread_all(r->fd, &c, 1, p);
	return c;
}

// FASTQ: first record start >= p: a line starting with '@' whose next-but-one line starts with '+'
bool next_fastq_record(btlbf_seqfile* r, uint64_t p, uint64_t* out)
{
	uint64_t ls;
	if (!next_line_start(r, p, &ls))
		return false;
	for (int tries = 0; tries < 16 && ls < r->size; tries++) {
		if (byte_at(r, ls) == '@') {
			uint64_t l1, l2;
			if (!next_line_start(r, ls + 1, &l1) || !next_line_start(r, l1 + 1, &l2))
				return false;
			if (l2 >= r->size || byte_at(r, l2) == '+') {
				*out = ls;
				return true;
			}
		}
		if (!next_line_start(r, ls + 1, &ls))
			return false;
	}
	*out = ls < r->size ? ls : r->size;
	return true;
}

// FASTA: the last <= overlap sequence bases before line start p (empty when the previous line is a header)
bool collect_tail(btlbf_seqfile* r, uint64_t p, std::string* tail)
{
	tail->clear();
	if (p == 0 || r->overlap == 0)
		return true;
	std::string rev;
	uint64_t end = p; // exclusive end of the bytes still to look at; byte end-1 is a '\n' (line start p)
	std::vector<char> buf;
	while (end > 0 && rev.size() < r->overlap) {
		// find the start of the line that ends at end-1
		uint64_t line_end = end - 1; // position of the terminating '\n'
		uint64_t ls = line_end;
		for (;;) {
			if (ls == 0)
				break;
			size_t n = (size_t)std::min<uint64_t>(65536, ls);
			buf.resize(n);
			if (!pread_all(r->fd, buf.data(), n, ls - n))
				return false;
			const char* q = (const char*)memrchr(buf.data(), '\n', n);
			if (q) {
				ls = ls - n + (uint64_t)(q - buf.data()) + 1;
				break;
			}
			ls -= n;
		}
		char first = byte_at(r, ls);
		if (first == '>')
			break; // the sequence starts right after this header: nothing more to prepend
		if (first != ';') {
			// append the line's bases in reverse
			uint64_t len = line_end - ls;
			uint64_t take = std::min<uint64_t>(len, (uint64_t)r->overlap - rev.size() + 2);
			buf.resize((size_t)take);
			if (take && !pread_all(r->fd, buf.data(), (size_t)take, line_end - take))
				return false;
			for (uint64_t i = take; i > 0 && rev.size() < r->overlap; i--) {
				char c = buf[(size_t)i - 1];
				if (c != '\r' && c != '\n')
					rev.push_back(c);
			}
		}
		end = ls;
	}
	tail->assign(rev.rbegin(), rev.rend());
	return true;
}

} // namespace

extern "C" int btlbf_seqfile_open(const char* path, unsigned overlap, int n_regions, int region, btlbf_seqfile** out)
{
	if (!path || !out || n_regions < 1 || region < 0 || region >= n_regions)
		return failf(BTLBF_ERR_ARG, "bad arguments to btlbf_seqfile_open");
	*out = nullptr;
	int fd = open(path, O_RDONLY);
	if (fd < 0)
		return failf(BTLBF_ERR_ARG, "cannot open '%s'", path);
	struct stat sb;
	if (fstat(fd, &sb) != 0) {
		close(fd);
		return failf(BTLBF_ERR_ARG, "cannot stat '%s'", path);
	}
	btlbf_seqfile* r = new btlbf_seqfile();
	r->fd = fd;
	r->size = (uint64_t)sb.st_size;
	r->overlap = overlap;
	r->path = path;
	char first = byte_at(r, 0);
	if (r->size && first != '>' && first != '@' && first != ';') {
		delete r;
		close(fd);
		return failf(BTLBF_ERR_ARG, "'%s' is neither FASTA ('>') nor FASTQ ('@')", path);
	}
	r->fastq = first == '@';
	uint64_t lo = r->size / (uint64_t)n_regions * (uint64_t)region;
	uint64_t hi = region + 1 == n_regions ? r->size : r->size / (uint64_t)n_regions * (uint64_t)(region + 1);
	bool ok = true;
	if (r->fastq) {
		ok = next_fastq_record(r, lo, &lo) && (hi >= r->size || next_fastq_record(r, hi, &hi));
	} else {
		ok = next_line_start(r, lo, &lo) && (hi >= r->size || next_line_start(r, hi, &hi));
		if (ok && lo < hi && lo > 0) {
			char c = byte_at(r, lo);
			if (c != '>' && c != ';') {
				// the region starts inside a sequence that began in an earlier region
				ok = collect_tail(r, lo, &r->tail);
				r->seq_open = true;
			}
		}
	}
	if (!ok) {
		delete r;
		close(fd);
		return failf(BTLBF_ERR_ARG, "read error in '%s'", path);
	}
	r->pos = lo;
	r->hi = hi;
	r->done = lo >= hi;
	*out = r;
	return BTLBF_OK;
}

extern "C" int btlbf_seqfile_close(btlbf_seqfile* r)
{
	if (r) {
		if (r->fd >= 0)
			close(r->fd);
		delete r;
	}
	return BTLBF_OK;
}

// Fills one flat batch: bases[0..*n_bases), offsets[0..*n_seqs] (pieces), *n_records = records that STARTED in
// this batch.  *done = 1 when the region is exhausted (the batch may still carry data).
extern "C" int btlbf_seqfile_next(btlbf_seqfile* r, char* bases, uint64_t cap_bases, uint64_t* offsets,
                                  uint64_t cap_seqs, uint64_t* n_bases, uint64_t* n_seqs, uint64_t* n_records, int* done)
{
	if (!r || !bases || !offsets || !n_bases || !n_seqs || !done || cap_seqs < 1 || cap_bases < (uint64_t)r->overlap + 1)
		return failf(BTLBF_ERR_ARG, "bad arguments to btlbf_seqfile_next");
	uint64_t nb = 0, ns = 0, nrec = 0;
	offsets[0] = 0;
	// pieces: offsets[ns] is the start of the piece being written; close_piece() seals it (empty pieces vanish)
	r->piece_open = false;
	auto close_piece = [&]() {
		if (r->piece_open && nb > offsets[ns])
			offsets[++ns] = nb;
		r->piece_open = false;
	};
	if (r->seq_open && !r->done) {
		// continuation of a sequence cut by the previous batch (or by the region boundary): its last k-1 bases first
		r->piece_open = true;
		memcpy(bases, r->tail.data(), r->tail.size());
		nb = r->tail.size();
	}
	bool full = false;
	while (!r->done && !full) {
		if (r->st == LINE_START && r->file_pos() >= r->hi) {
			r->done = true;
			break;
		}
		if (!r->fill()) {
			if (r->file_pos() >= r->size) {
				r->done = true;
				break;
			}
			return failf(BTLBF_ERR_ARG, "read error in '%s'", r->path.c_str());
		}
		const char* p = r->raw.data() + r->raw_off;
		const size_t avail = r->raw_len - r->raw_off;
		if (r->st == LINE_START) {
			const char c = *p;
			const bool header = r->fastq ? r->fq_line == 0 : c == '>';
			const bool sequence = r->fastq ? r->fq_line == 1 : (c != '>' && c != ';');
			if (header) {
				close_piece();
				r->seq_open = false;
				if (ns + 1 >= cap_seqs) { // no room for another piece: this record starts the next batch
					full = true;
					break;
				}
				nrec++;
				r->st = IN_HEADER;
			} else if (sequence) {
				r->piece_open = true;
				r->seq_open = true;
				r->st = IN_SEQ;
			} else
				r->st = IN_SKIP;
		}
		const char* nl = (const char*)memchr(p, '\n', avail);
		const size_t len = nl ? (size_t)(nl - p) : avail; // bytes of this line in the block, newline excluded
		if (r->st == IN_SEQ) {
			const size_t room = (size_t)(cap_bases - nb);
			const size_t take = len < room ? len : room;
			memcpy(bases + nb, p, take);
			nb += take;
			r->raw_off += take;
			if (take < len) { // the batch is full in the middle of a line: the rest goes to the next batch
				full = true;
				break;
			}
			if (nl && nb > offsets[ns] && bases[nb - 1] == '\r') // CRLF files
				nb--;
			if (nb == cap_bases)
				full = true;
		} else
			r->raw_off += len;
		if (nl) {
			r->raw_off++; // the newline
			if (r->fastq) {
				if (r->fq_line == 1)
					r->seq_open = false; // four-line FASTQ records: the sequence is one line
				r->fq_line = (r->fq_line + 1) & 3;
			}
			r->st = LINE_START;
		}
		if (r->raw_off >= r->raw_len)
			r->advance_block();
	}
	// the tail of a sequence that continues: start of its next piece
	if (r->seq_open && r->piece_open && !r->done) {
		uint64_t plen = nb - offsets[ns];
		uint64_t t = plen < r->overlap ? plen : r->overlap;
		r->tail.assign(bases + nb - t, (size_t)t);
	} else if (!r->seq_open)
		r->tail.clear();
	close_piece();
	*n_bases = nb;
	*n_seqs = ns;
	if (n_records)
		*n_records = nrec;
	*done = r->done ? 1 : 0;
	return BTLBF_OK;
}

// ---------------------------------------------------------------- file -> filter
namespace {

struct Batch
{
	char* bases = nullptr; // pinned
	uint64_t* offsets = nullptr;
	uint64_t n_bases = 0, n_seqs = 0;
	uint64_t* counts = nullptr; // pinned {n_kmers, n_hits}
};

// Pinned staging buffers are expensive to create (~0.4 s per GB): they are kept for the life of the process and
// handed out to one file call at a time (btlbf_ingest_release frees them).
struct PinnedCache
{
	std::mutex mu;
	std::vector<char*> bases; // each `cap` bytes
	std::vector<uint64_t*> offsets; // each kIngestBatch / 16 + 1 entries
	uint64_t* counts = nullptr;
	size_t counts_slots = 0;
	bool busy = false;
} g_pinned;
constexpr uint64_t kIngestBatch = (uint64_t)32 << 20; // bases per batch

struct Pool
{
	std::mutex mu;
	std::condition_variable cv;
	std::deque<Batch*> free_list, ready;
	int producers = 0;
	int error = 0;
	std::string error_msg;
};

int file_op(btlbf_filter* f, const char* path, bool query, int threads, uint64_t* n_seqs, uint64_t* n_kmers,
            uint64_t* n_hits)
{
	if (!f || !path)
		return failf(BTLBF_ERR_ARG, "null argument");
	int kind = 0;
	unsigned k = 0;
	int rc = btlbf_filter_info(f, &kind, nullptr, nullptr, nullptr, &k, nullptr);
	if (rc != BTLBF_OK)
		return rc;
	// order matters for the counting insert (the reference result is the file-order loop): one reader
	if (!query && kind != BTLBF_BLOOM)
		threads = 1;
	if (threads < 1)
		threads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 8u);
	struct stat sb;
	if (stat(path, &sb) != 0)
		return failf(BTLBF_ERR_ARG, "cannot open '%s'", path);
	if ((uint64_t)sb.st_size < ((uint64_t)threads << 20))
		threads = 1;
	const uint64_t cap = kIngestBatch;
	const uint64_t cap_seqs = cap / 16;
	const int n_batches = threads + 6;       // being filled + in flight (kTickets) + slack
	std::vector<Batch> store((size_t)n_batches);
	Pool pool;
	{
		std::unique_lock<std::mutex> lk(g_pinned.mu);
		if (g_pinned.busy)
			return failf(BTLBF_ERR_STATE, "another file call is in progress in this process");
		if (g_pinned.counts_slots < (size_t)n_batches) {
			if (g_pinned.counts)
				cudaFreeHost(g_pinned.counts);
			g_pinned.counts = nullptr;
			g_pinned.counts_slots = 0;
			if (cudaHostAlloc((void**)&g_pinned.counts, (size_t)n_batches * 16, cudaHostAllocDefault) != cudaSuccess)
				return failf(BTLBF_ERR_NOMEM, "pinned allocation failed");
			g_pinned.counts_slots = (size_t)n_batches;
		}
		while (g_pinned.bases.size() < (size_t)n_batches) {
			char* b = nullptr;
			if (cudaHostAlloc((void**)&b, cap, cudaHostAllocDefault) != cudaSuccess)
				return failf(BTLBF_ERR_NOMEM, "pinned allocation of %llu bytes failed", (unsigned long long)cap);
			g_pinned.bases.push_back(b);
		}
		while (g_pinned.offsets.size() < (size_t)n_batches) {
			uint64_t* o = (uint64_t*)malloc((cap_seqs + 1) * sizeof(uint64_t));
			if (!o)
				return failf(BTLBF_ERR_NOMEM, "out of host memory");
			g_pinned.offsets.push_back(o);
		}
		g_pinned.busy = true;
	}
	auto cleanup = [&]() {
		std::unique_lock<std::mutex> lk(g_pinned.mu);
		g_pinned.busy = false;
	};
	for (int i = 0; i < n_batches; i++) {
		Batch& b = store[(size_t)i];
		b.bases = g_pinned.bases[(size_t)i];
		b.offsets = g_pinned.offsets[(size_t)i];
		b.counts = g_pinned.counts + 2 * i;
		b.counts[0] = b.counts[1] = 0;
		pool.free_list.push_back(&b);
	}
	std::vector<btlbf_seqfile*> readers((size_t)threads, nullptr);
	for (int t = 0; t < threads; t++) {
		rc = btlbf_seqfile_open(path, k ? k - 1 : 0, threads, t, &readers[(size_t)t]);
		if (rc != BTLBF_OK) {
			for (btlbf_seqfile* r : readers)
				btlbf_seqfile_close(r);
			cleanup();
			return rc;
		}
	}
	std::vector<uint64_t> records((size_t)threads, 0);
	pool.producers = threads;
	auto produce = [&](int t) {
		btlbf_seqfile* r = readers[(size_t)t];
		int done = 0;
		while (!done) {
			Batch* b;
			{
				std::unique_lock<std::mutex> lk(pool.mu);
				pool.cv.wait(lk, [&] { return !pool.free_list.empty() || pool.error; });
				if (pool.error)
					break;
				b = pool.free_list.front();
				pool.free_list.pop_front();
			}
			uint64_t nrec = 0;
			int e = btlbf_seqfile_next(r, b->bases, cap, b->offsets, cap_seqs, &b->n_bases, &b->n_seqs, &nrec, &done);
			records[(size_t)t] += nrec;
			std::unique_lock<std::mutex> lk(pool.mu);
			if (e != BTLBF_OK) {
				pool.error = e;
				pool.error_msg = btlbf_last_error();
				pool.free_list.push_back(b);
				pool.cv.notify_all();
				break;
			}
			if (b->n_seqs)
				pool.ready.push_back(b);
			else
				pool.free_list.push_back(b);
			pool.cv.notify_all();
		}
		std::unique_lock<std::mutex> lk(pool.mu);
		pool.producers--;
		pool.cv.notify_all();
	};
	std::vector<std::thread> workers;
	for (int t = 0; t < threads; t++)
		workers.emplace_back(produce, t);

	// consumer: this thread owns the filter handle
	std::deque<Batch*> in_flight;
	uint64_t tot_k = 0, tot_h = 0;
	auto retire = [&](Batch* b) {
		tot_k += b->counts[0];
		tot_h += b->counts[1];
		std::unique_lock<std::mutex> lk(pool.mu);
		pool.free_list.push_back(b);
		pool.cv.notify_all();
	};
	int err = BTLBF_OK;
	for (;;) {
		Batch* b = nullptr;
		{
			std::unique_lock<std::mutex> lk(pool.mu);
			pool.cv.wait(lk, [&] { return !pool.ready.empty() || pool.producers == 0 || pool.error; });
			if (pool.error) {
				err = pool.error;
				break;
			}
			if (pool.ready.empty())
				break; // all producers are done
			b = pool.ready.front();
			pool.ready.pop_front();
		}
		if (query)
			err = btlbf_contains_seqs_async(f, b->bases, b->offsets, b->n_seqs, nullptr, nullptr, b->counts);
		else
			err = btlbf_insert_seqs_async(f, b->bases, b->offsets, b->n_seqs, b->counts);
		if (err != BTLBF_OK) {
			std::unique_lock<std::mutex> lk(pool.mu);
			pool.error = err;
			pool.error_msg = btlbf_last_error();
			pool.cv.notify_all();
			break;
		}
		in_flight.push_back(b);
		// a call's buffers are free once four later calls have been queued (the ticket it used was waited for)
		while (in_flight.size() > 4) {
			retire(in_flight.front());
			in_flight.pop_front();
		}
	}
	for (std::thread& w : workers)
		w.join();
	int sync_rc = BTLBF_OK;
	{
		// everything queued completes before the pinned buffers go away
		btlbf_ctx* ctx = nullptr;
		sync_rc = btlbf_filter_ctx(f, &ctx);
		if (sync_rc == BTLBF_OK)
			sync_rc = btlbf_ctx_sync(ctx);
	}
	while (!in_flight.empty()) {
		retire(in_flight.front());
		in_flight.pop_front();
	}
	uint64_t nrec = 0;
	for (uint64_t x : records)
		nrec += x;
	for (btlbf_seqfile* r : readers)
		btlbf_seqfile_close(r);
	cleanup();
	if (err != BTLBF_OK)
		return btlbf_set_error(err, pool.error_msg.empty() ? btlbf_last_error() : pool.error_msg.c_str());
	if (sync_rc != BTLBF_OK)
		return sync_rc;
	if (n_seqs) *n_seqs = nrec;
	if (n_kmers) *n_kmers = tot_k;
	if (n_hits) *n_hits = tot_h;
	return BTLBF_OK;
}

} // namespace

extern "C" int btlbf_ingest_release(void)
{
	std::unique_lock<std::mutex> lk(g_pinned.mu);
	if (g_pinned.busy)
		return failf(BTLBF_ERR_STATE, "a file call is in progress");
	for (char* b : g_pinned.bases)
		cudaFreeHost(b);
	g_pinned.bases.clear();
	for (uint64_t* o : g_pinned.offsets)
		free(o);
	g_pinned.offsets.clear();
	if (g_pinned.counts)
		cudaFreeHost(g_pinned.counts);
	g_pinned.counts = nullptr;
	g_pinned.counts_slots = 0;
	return BTLBF_OK;
}

extern "C" int btlbf_insert_file(btlbf_filter* f, const char* path, int threads, uint64_t* n_seqs, uint64_t* n_kmers)
{
	return file_op(f, path, false, threads, n_seqs, n_kmers, nullptr);
}

extern "C" int btlbf_query_file(btlbf_filter* f, const char* path, int threads, uint64_t* n_seqs, uint64_t* n_kmers,
                                uint64_t* n_hits)
{
	return file_op(f, path, true, threads, n_seqs, n_kmers, n_hits);
}
