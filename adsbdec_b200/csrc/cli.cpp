// cli.cpp - command line with the option letters and observable behaviour of the reference's
// main.c:54-106 / output.c:264-344 for the `-f FILE` path, on top of the C ABI.
//
//   adsbdec_b200 [-a] [-m] [-b] [-g N] -f FILE [-s addr[:port]] [-l addr[:port]]
//
// Kept from the reference: getopt string "cf:s:l:g:aemb" with -c / -e / unknown -> usage, exit 1
// (main.c:60, :85-87); -m / -b last one wins (main.c:79-84); -g parsed and ignored with -f; frames
// written to stdout and flushed one by one (output.c:331-332) or to a TCP server (-s); the stats
// block on stderr at exit (valid.c:84-101); a file that cannot be opened prints stats and exits 0
// (air.c:225-228).  Deliberately NOT kept: stdout beast output truncated by strlen() (output.c:331
// - we write the full packet) and the frames the reference loses in its EOF race (output.c:287).
// Extra: -D n selects the CUDA device, -G forces the generic kernels, -P accepts address/parity frames
// (DF4/5/20/21), -R 1|2 switches on the error
// repair extension (repair.h; the reference has none, so the default output is the reference's).
#include <arpa/inet.h>
#include <netdb.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>

#include <string>

#include "../../include/adsbdec_b200.h"

namespace {

int g_format = 0;  // 0 avr, 1 avrmlat, 2 beast (output.c:40)
int g_sock = -1;
bool g_network = false;  // -s / -l given: outmode 1 / 2 of the reference (output.c:38)

void usage()
{
	printf("adsbdec_b200 - B200-native ADS-B decoder, drop-in for `adsbdec -f`\n\n");
	printf("usage : adsbdec_b200 [-a] [-m] [-b] [-g 0-21] -f filename [-s addr[:port]] [-l addr[:port]]\n\n");
	printf("Options :\n");
	printf("\t-a : decode DF18 too\n");
	printf("\t-m : output avrmlat format (ie : with 12Mhz timestamp)\n");
	printf("\t-b : output binary beast format\n");
	printf("\t-g 0-21 : accepted for compatibility, ignored with -f\n");
	printf("\t-f : input file (airspy raw real samples, 16 bits, 20 MS/s)\n");
	printf("\t-s addr[:port] : send output via TCP to server at address addr:port (default port : 30001)\n");
	printf("\t-l addr[:port] : listen on addr:port and send output via TCP to the first client (default port : 30002)\n");
	printf("\t-D n : CUDA device ordinal\n");
	printf("\t-R 1|2 : extension, not in adsbdec : repair 1 (2) flipped bits by CRC syndrome\n");
	printf("\t-P : extension, not in adsbdec : also DF4/5/20/21 of aircraft announced earlier (address/parity)\n");
	printf("\t-I : extension, not in adsbdec : the file holds signed 16-bit I/Q pairs at 10 MS/s\n");
}

// -s: connect to a server; -l: bind, wait for ONE client ("listening" / "connected" on stderr), then
// write to it exactly like -s (output.c:59-157; default ports 30001 / 30002).
int open_socket(const char *spec, bool listener)
{
	// "addr[:port]" or "[v6addr]:port"
	std::string host(spec), port(listener ? "30002" : "30001");
	if (!host.empty() && host[0] == '[') {
		size_t e = host.find(']');
		if (e != std::string::npos) {
			if (e + 1 < host.size() && host[e + 1] == ':')
				port = host.substr(e + 2);
			host = host.substr(1, e - 1);
		}
	} else {
		size_t c = host.rfind(':');
		if (c != std::string::npos) {
			port = host.substr(c + 1);
			host = host.substr(0, c);
		}
	}
	addrinfo hints{}, *res = nullptr;
	hints.ai_socktype = SOCK_STREAM;
	if (getaddrinfo(host.c_str(), port.c_str(), &hints, &res) != 0)
		return -1;
	int fd = -1;
	for (addrinfo *p = res; p; p = p->ai_next) {
		fd = socket(p->ai_family, p->ai_socktype, p->ai_protocol);
		if (fd < 0)
			continue;
		if (!listener) {
			if (connect(fd, p->ai_addr, p->ai_addrlen) == 0)
				break;
		} else {
			int one = 1;
			setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
			if (bind(fd, p->ai_addr, p->ai_addrlen) == 0 && listen(fd, 1) == 0) {
				fprintf(stderr, "listening\n");
				const int client = accept(fd, nullptr, nullptr);
				close(fd);
				fd = client;
				if (fd >= 0)
					break;
				continue;
			}
		}
		close(fd);
		fd = -1;
	}
	freeaddrinfo(res);
	return fd;
}

void sink(const uint8_t *frame, int len, uint64_t ts, uint32_t pw, void *)
{
	char pkt[256];
	int n = adsbdec_b200_format_packet(g_format, frame, len, ts, pw, pkt);
	if (n <= 0)
		return;
	if (g_network) {
		// -s / -l: frames only ever go to the socket.  After a write error the reference drops its queue
		// and reconnects (output.c:320-327, :290-299); with a file as input there is nothing to wait for,
		// so the remaining frames are dropped - never printed on stdout (binary beast with -b).
		if (g_sock < 0)
			return;
		const char *p = pkt;
		while (n > 0) {
			ssize_t w = write(g_sock, p, (size_t)n);
			if (w <= 0) {
				fprintf(stderr, "disconnected\n");
				close(g_sock);
				g_sock = -1;
				return;
			}
			p += w;
			n -= (int)w;
		}
	} else {
		fwrite(pkt, (size_t)n, 1, stdout);
		fflush(stdout);
	}
}

void print_stats(adsbdec_b200_t *h)
{
	uint32_t t[3] = {0, 0, 0}, k[3] = {0, 0, 0};
	if (h)
		adsbdec_b200_stats(h, t, k);
	// valid.c:88-100 (the reference adds an uninitialised tot_fi to Total; we print the sum of Ok)
	fprintf(stderr, "\t%10d\t%10d\t%10d\n", 11, 17, 18);
	fprintf(stderr, "Try :\t%10d\t%10d\t%10d\n", (int)t[0], (int)t[1], (int)t[2]);
	fprintf(stderr, "Ok :\t%10d\t%10d\t%10d\n", (int)k[0], (int)k[1], (int)k[2]);
	fprintf(stderr, "Total :\t%10d\n", (int)(k[0] + k[1] + k[2]));
}

}  // namespace

int main(int argc, char **argv)
{
	const char *filename = nullptr, *addr = nullptr;
	bool listener = false;
	adsbdec_b200_config_t cfg;
	adsbdec_b200_default_config(&cfg);
	int c;
	while ((c = getopt(argc, argv, "cf:s:l:g:aembD:GR:PI")) != EOF) {
		switch (c) {
		case 'f': filename = optarg; break;
		case 's': addr = optarg; listener = false; break;  // main.c: outmode 1
		case 'l': addr = optarg; listener = true; break;   // main.c: outmode 2
		case 'g': (void)atoi(optarg); break;
		case 'a': cfg.accept_df18 = 1; break;
		case 'm': g_format = 1; break;
		case 'b': g_format = 2; break;
		case 'D': cfg.device = atoi(optarg); break;
		case 'G': cfg.kernel = 1; break;
		case 'R': cfg.reserved[3] = (cfg.reserved[3] & 0x100u) | (uint64_t)(atoi(optarg) & 0xFF); break;
		case 'P': cfg.reserved[3] |= 0x100u; break;
		case 'I': cfg.reserved[3] |= 0x200u; break;
		default:  // -c, -e and anything unknown
			usage();
			return 1;
		}
	}
	signal(SIGPIPE, SIG_IGN);
	if (!filename) {
		fprintf(stderr, "adsbdec_b200: only the -f file path is implemented (no SDR input)\n");
		usage();
		return 1;
	}
	if (addr) {
		g_network = true;
		g_sock = open_socket(addr, listener);
		if (g_sock < 0) {
			fprintf(stderr, "adsbdec_b200: cannot %s %s\n", listener ? "listen on" : "connect to", addr);
			return 255;
		}
		fprintf(stderr, "connected\n");
	}
	adsbdec_b200_t *h = nullptr;
	int rc = adsbdec_b200_open(&cfg, &h);
	if (rc) {
		fprintf(stderr, "adsbdec_b200: %s\n", adsbdec_b200_last_error());
		return 2;
	}
	adsbdec_b200_set_sink(h, sink, nullptr);
	rc = adsbdec_b200_decode_file(h, filename);
	int status = 0;
	if (rc && rc != ADSB_B200_EIO) {
		fprintf(stderr, "adsbdec_b200: %s\n", adsbdec_b200_last_error());
		status = 2;
	}
	print_stats(h);
	adsbdec_b200_close(h);
	if (g_sock >= 0)
		close(g_sock);
	return status;
}
