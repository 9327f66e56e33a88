// bench_blight_b200.cpp — the reference's command-line driver shape (bench_blight.cpp:37-108: -g -q -k -m -n -s -t -b,
// self-query when -q is omitted) on top of the drop-in class.
#include <unistd.h>

#include <chrono>
#include <iostream>
#include <string>

#include "kmer_set_light.hpp"

int main(int argc, char** argv) {
	std::string input, query;
	unsigned k = 31, m1 = 9, m2 = 17, m3 = 6, c = 1, bit = 6;
	int ch;
	while ((ch = getopt(argc, argv, "g:q:k:m:n:s:t:b:")) != -1) {
		switch (ch) {
			case 'q': query = optarg; break;
			case 'g': input = optarg; break;
			case 'k': k = std::stoi(optarg); break;
			case 'm': m1 = std::stoi(optarg); break;
			case 'n': m2 = std::stoi(optarg); break;
			case 's': m3 = std::stoi(optarg); break;
			case 't': c = std::stoi(optarg); break;
			case 'b': bit = std::stoi(optarg); break;
		}
	}
	if (query.empty()) query = input;
	if (input.empty()) input = query;
	if (input.empty() || k == 0) {
		std::cout << "Mandatory arguments\n\t-g graph file\n\t-q query file\n\t-k k value used for graph (" << k << ")\n\n"
		          << "Performances arguments\n\t-m minimizer size (" << m1 << ")\n\t-n to create 2^n mphf (" << m2
		          << ")\n\t-s to use 2^s files (" << m3 << ")\n\t-t core used (" << c << ")\n\t-b bit saved to encode positions (" << bit << ")\n";
		return 0;
	}
	try {
		using clk = std::chrono::high_resolution_clock;
		kmer_Set_Light ksl(k, m1, m2, m3, c, bit);
		auto t0 = clk::now();
		ksl.construct_index(input);
		auto t1 = clk::now();
		std::cout << "The whole indexing took me " << std::chrono::duration<double>(t1 - t0).count() << " seconds.\n";
		std::cout << "Kmer in graph: " << ksl.number_kmer << "\nSuper Kmer in graph: " << ksl.number_super_kmer << "\n";
		auto r = ksl.file_query(query);
		auto t2 = clk::now();
		std::cout << "-----------------------QUERY RECAP 2----------------------------\n";
		std::cout << "Good kmer: " << r.first << "\nErroneous kmers: " << r.second << "\nQuery performed: " << ksl.number_query.load() << "\n";
		std::cout << "The whole QUERY took me " << std::chrono::duration<double>(t2 - t1).count() << " seconds.\n";
	} catch (const std::exception& e) {
		std::cerr << "error: " << e.what() << "\n";
		return 1;
	}
	return 0;
}
