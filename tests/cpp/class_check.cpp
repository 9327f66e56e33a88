// Exercises the drop-in class (blight_b200/csrc/kmer_set_light.hpp) the way the reference's callers use kmer_Set_Light
// (bench_blight.cpp:99-102, the snippet applications): exceptions of the constructor and of the query calls, and —
// with a GPU — index construction, per-sequence and per-k-mer queries, file_query. Prints one "key value" line per fact.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>

#include "kmer_set_light.hpp"

template <class E, class F>
static bool throws(F f) {
	try { f(); } catch (const E&) { return true; } catch (...) { return false; }
	return false;
}

int main(int argc, char** argv) {
	const std::string mode = argc > 1 ? argv[1] : "cpu";
	const std::string fasta = argc > 2 ? argv[2] : "";
	// blight.h:75-92: even minimizer length, too many MPHFs, ... -> std::invalid_argument
	std::cout << "invalid_even_m " << throws<std::invalid_argument>([] { kmer_Set_Light x(31, 8, 5, 3, 1, 6); }) << "\n";
	std::cout << "invalid_big_n " << throws<std::invalid_argument>([] { kmer_Set_Light x(31, 7, 14, 3, 1, 6); }) << "\n";
	std::cout << "valid_params " << !throws<std::exception>([] { kmer_Set_Light x(31, 7, 5, 3, 1, 6); }) << "\n";
	{
		kmer_Set_Light ksl(31, 7, 5, 3, 1, 6);
		std::cout << "query_before_index " << throws<std::runtime_error>([&] { ksl.query_sequence_hash(std::string(40, 'A')); }) << "\n";
		std::cout << "missing_file " << throws<std::runtime_error>([&] { ksl.construct_index("/nonexistent/unitigs.fa"); }) << "\n";
	}
	if (mode == "cpu") {
		// no GPU here: construct_index must fail loudly (no CPU fallback), never answer queries
		kmer_Set_Light ksl(31, 7, 5, 3, 1, 6);
		std::cout << "no_device_is_an_error " << throws<std::runtime_error>([&] { ksl.construct_index(fasta); }) << "\n";
		return 0;
	}
	kmer_Set_Light ksl(31, 7, 5, 3, 4, 6);
	if (argc > 4) {
		// several devices from one process: argv[3] = replica | partition, argv[4..] = device ordinals (repeats allowed)
		std::vector<int> devs;
		for (int i = 4; i < argc; i++) devs.push_back(std::stoi(argv[i]));
		ksl.use_devices(devs, std::string(argv[3]) == "partition" ? BLIGHT_COMM_PARTITION : BLIGHT_COMM_REPLICA);
	}
	ksl.construct_index(fasta);
	std::cout << "number_kmer " << ksl.number_kmer << "\nnumber_super_kmer " << ksl.number_super_kmer << "\n";
	std::ifstream in(fasta);
	std::string header, seq;
	size_t unitig = 0;
	while (std::getline(in, header) && std::getline(in, seq) && seq.size() < 250) unitig++;  // first unitig long enough
	std::cout << "unitig " << unitig << "\n";
	const std::string read = seq.substr(100, 150);
	const auto ids = ksl.query_sequence_hash(read);
	std::cout << "ids_n " << ids.size() << "\nids_first " << ids.front() << "\nids_last " << ids.back() << "\n";
	const auto gb = ksl.query_sequence_bool(read);
	std::cout << "bool " << gb.first << " " << gb.second << "\n";
	std::cout << "short_read_empty " << ksl.query_sequence_hash(read.substr(0, 30)).empty() << "\n";
	std::string bad = read;
	bad[75] = 'N';
	std::cout << "domain_error_on_N " << throws<std::domain_error>([&] { ksl.query_sequence_hash(bad); }) << "\n";
	// canonical k-mer of the read's first 31 bases, computed like nuc2int / min_k (kmer.h:56-98, blight.cpp:86-91)
	auto code = [](char c) { return (uint64_t)((c >> 1) & 3); };
	uint64_t f = 0, r = 0;
	for (int i = 0; i < 31; i++) { f = (f << 2) | code(read[i]); r = (r >> 2) | ((code(read[i]) ^ 2) << 60); }
	const uint64_t canon = f < r ? f : r;
	std::cout << "kmer_hash_equals_first_id " << (ksl.query_kmer_hash(canon) == ids.front()) << "\nkmer_bool " << ksl.query_kmer_bool(canon) << "\n";
	std::cout << "absent_kmer " << ksl.query_kmer_hash(0x0123456789ABCDEull & ((1ull << 62) - 1)) << "\n";
	const auto fq = ksl.file_query(fasta);
	std::cout << "file_query " << fq.first << " " << fq.second << "\nnumber_query " << ksl.number_query.load() << "\n";
	return 0;
}
