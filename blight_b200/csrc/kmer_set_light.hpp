// kmer_set_light.hpp — the drop-in C++ surface: a class with the reference's public query interface
// (kmer_Set_Light, blight.h:15-136) whose work is done by the sm_100a library behind include/blight_b200.h.
//
// Same constructor arguments and exceptions (std::invalid_argument, blight.h:75-92), same method names,
// argument meaning and results: construct_index (blight.h:134), file_query (:117), query_sequence_bool (:129),
// query_sequence_hash (:133), query_kmer_bool (:128), query_kmer_hash (:131) and the public counters number_kmer,
// number_super_kmer, number_query (:52-56).  Non-ACGT input raises std::domain_error like nuc2int (kmer.h:68);
// an unreadable file raises std::runtime_error (blight.cpp:188-189).  Like the reference object it is
// non-copyable; queries may be issued concurrently from several host threads.
// Ours: use_devices({0,1,..,7}, BLIGHT_COMM_REPLICA | BLIGHT_COMM_PARTITION) before construct_index / load_index spreads
// the same object over several GPUs of the box (csrc/comm.cu); every method keeps its results.
#pragma once
#include <atomic>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "blight_b200.h"

using kmer_t = uint64_t;

class kmer_Set_Light {
public:
	size_t number_kmer = 0;
	size_t number_super_kmer = 0;
	std::atomic<size_t> number_query{0};

	kmer_Set_Light(unsigned k, unsigned minimizer_length, unsigned log2_mphfs_number, unsigned log2_superbuckets_number,
	               unsigned cores_number, unsigned bits_to_save, int device = 0)
	    : _k(k), _m(minimizer_length), _n(log2_mphfs_number), _s(log2_superbuckets_number), _cores(cores_number), _b(bits_to_save), _device(device) {
		if (blight_check_params(_k, _m, _n, _s, _b) != BLIGHT_OK) throw std::invalid_argument(blight_last_error());
	}
	kmer_Set_Light(const kmer_Set_Light&) = delete;
	kmer_Set_Light& operator=(const kmer_Set_Light&) = delete;
	~kmer_Set_Light() { drop(); }

	// ours: several GPUs of the box (call before construct_index / load_index)
	void use_devices(const std::vector<int>& devices, int comm_mode) {
		if (devices.empty()) throw std::invalid_argument("use_devices: no device");
		_devices = devices;
		_comm_mode = comm_mode;
	}

	void construct_index(const std::string& input_file) {
		drop();
		raise(blight_flat_build_file(input_file.c_str(), _k, _m, _n, _s, _b, _cores, &_flat));
		adopt();
	}

	// ours: load an index saved with save_index() (or exported from a reference object)
	void load_index(const std::string& blob) {
		drop();
		raise(blight_flat_load(blob.c_str(), &_flat));
		adopt();
	}
	void save_index(const std::string& blob) const { raise(blight_flat_save(_flat, blob.c_str())); }

	// Returns (Good kmer, Erroneous kmers) and adds to number_query; the reference prints them (blight.cpp:792-795).
	std::pair<uint64_t, uint64_t> file_query(const std::string& query_file) {
		uint64_t ctr[BLIGHT_N_CTR];
		if (_comm) raise(blight_comm_query_file_host(_comm, query_file.c_str(), ctr));
		else raise(blight_query_file_host(need(), query_file.c_str(), ctr));
		number_query += ctr[BLIGHT_CTR_QUERIES];
		return {ctr[BLIGHT_CTR_FOUND], ctr[BLIGHT_CTR_NOT_FOUND]};
	}

	std::vector<int64_t> query_sequence_hash(const std::string& query) {
		std::vector<int64_t> res(query.size() >= _k ? query.size() - _k + 1 : 0);
		uint64_t n = 0;
		if (_comm) raise(blight_comm_query_sequence_host(_comm, query.data(), query.size(), res.data(), &n));
		else raise(blight_query_sequence_host(need(), query.data(), query.size(), res.data(), &n));
		number_query += n;
		return res;
	}

	std::pair<uint32_t, uint32_t> query_sequence_bool(const std::string& query) {
		uint64_t good = 0, bad = 0;
		if (_comm) {
			if (query.size() >= _k) {
				const uint64_t off[2] = {0, query.size()};
				uint64_t ctr[BLIGHT_N_CTR];
				raise(blight_comm_query_reads_host(_comm, query.data(), off, 1, nullptr, ctr));
				good = ctr[BLIGHT_CTR_FOUND]; bad = ctr[BLIGHT_CTR_NOT_FOUND];
			}
		} else {
			raise(blight_query_sequence_bool_host(need(), query.data(), query.size(), &good, &bad));
		}
		number_query += good + bad;
		return {(uint32_t)good, (uint32_t)bad};
	}

	int64_t query_kmer_hash(kmer_t canon) {
		int64_t id = -1;
		if (_comm) {
			// the k-mer spelled out: the devices route it by its minimizer like any other query
			std::string s(_k, 'A');
			for (unsigned i = 0; i < _k; i++) s[i] = "ACTG"[(canon >> (2 * (_k - 1 - i))) & 3];
			uint64_t n = 0;
			raise(blight_comm_query_sequence_host(_comm, s.data(), s.size(), &id, &n));
		} else {
			raise(blight_query_kmers_host(need(), &canon, 1, &id));
		}
		number_query += 1;
		return id;
	}
	bool query_kmer_bool(kmer_t canon) { return query_kmer_hash(canon) >= 0; }

	// ours: batched forms of the two calls above
	std::vector<int64_t> query_kmers_hash(const std::vector<kmer_t>& canon) {
		std::vector<int64_t> ids(canon.size());
		if (_comm) {
			std::string s(canon.size() * _k, 'A');
			std::vector<uint64_t> off(canon.size() + 1);
			for (size_t j = 0; j < canon.size(); j++) {
				off[j] = j * _k;
				for (unsigned i = 0; i < _k; i++) s[j * _k + i] = "ACTG"[(canon[j] >> (2 * (_k - 1 - i))) & 3];
			}
			off[canon.size()] = canon.size() * _k;
			uint64_t ctr[BLIGHT_N_CTR];
			raise(blight_comm_query_reads_host(_comm, s.data(), off.data(), canon.size(), ids.data(), ctr));
		} else {
			raise(blight_query_kmers_host(need(), canon.data(), canon.size(), ids.data()));
		}
		number_query += canon.size();
		return ids;
	}

	const blight_index* device_index() const { return _idx; }

private:
	static void raise(int rc) {
		if (rc == BLIGHT_OK) return;
		const std::string msg = blight_last_error();
		if (rc == BLIGHT_ERR_INVALID_ARG) throw std::invalid_argument(msg);
		if (rc == BLIGHT_ERR_INVALID_BASE) throw std::domain_error(msg);
		throw std::runtime_error(msg);
	}
	const blight_index* need() const {
		if (!_idx) throw std::runtime_error("kmer_Set_Light: construct_index() has not been called");
		return _idx;
	}
	void drop() {
		blight_comm_free(_comm); _comm = nullptr;
		blight_index_free(_idx); _idx = nullptr;
		blight_flat_free(_flat); _flat = nullptr;
	}
	void adopt() {
		blight_info info;
		raise(blight_flat_info(_flat, &info));
		number_kmer = info.number_kmer;
		number_super_kmer = info.number_super_kmer;
		if (_devices.empty()) raise(blight_index_upload(_flat, _device, &_idx));
		else raise(blight_comm_init(_flat, _devices.data(), (uint32_t)_devices.size(), _comm_mode, nullptr, &_comm));
	}

	const unsigned _k, _m, _n, _s, _cores, _b;
	const int _device;
	blight_flat* _flat = nullptr;
	blight_index* _idx = nullptr;
	std::vector<int> _devices;
	int _comm_mode = BLIGHT_COMM_REPLICA;
	blight_comm* _comm = nullptr;
};
