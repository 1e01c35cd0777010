// setup.h -- host-side Vorbis header parsing and device table construction.
// Follows the behaviour of StreamDecoder.LoadStreamHeader/LoadBooks (StreamDecoder.cs:213-355),
// Codebook (Codebook.cs:21-298), Floor1 ctor (Floor1.cs:39-155), Residue0 ctor
// (Residue0.cs:25-115), Mapping ctor (Mapping.cs:19-95), Mode ctor (Mode.cs:14-28); the tables
// it emits are laid out for the GPU decoder (vpz_dev.h), not for the reference's CPU decoder.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "vpz_dev.h"

namespace vpz {

struct IdHeader {
  int channels = 0, sample_rate = 0;
  int br_upper = 0, br_nominal = 0, br_lower = 0;
  int size0 = 0, size1 = 0;
};

struct PacketGeom {          // PacketInfo.cs + what the host needs to place the packet
  bool valid = false;        // audio packet with a usable mode (StreamDecoder.cs:728-741)
  bool bad_mode = false;     // mode index >= modes: reference throws InvalidDataException
  bool long_block = false, prev_flag = true, next_flag = true;
  int mode = 0;
  int block_size = 0;
  int length = 0;            // overlap length of the left window
  bool left_use_size1 = false;
  int left_start = 0, left_end = 0, right_start = 0, right_end = 0;
};

class Setup {
 public:
  IdHeader id;
  std::vector<uint32_t> blob;   // device image, starts with VpzSetupHdr
  std::vector<VpzMode> modes;
  int mode_bits = 0;
  uint64_t hash = 0;            // content hash of (id packet, setup packet) for de-duplication
  int max_codeword_bits = 0;
  int n_residues_hdr = 0;       // residues in the setup header (the device image holds instances, vpz_dev.h)
  std::string error;

  // Returns 0 or a negative VPZ_E_* code (include/vpz.h); `error` holds the reason.
  int parse(const uint8_t* id_pkt, size_t id_len, const uint8_t* setup_pkt, size_t setup_len, int l1_bits);

  // Mode.GetPacketInfo (Mode.cs:30-66) + the packet-type / mode checks of
  // StreamDecoder.DecodeNextPacket (StreamDecoder.cs:728-741), from the packet's first bytes.
  PacketGeom packet_geometry(const uint8_t* pkt, size_t len) const;

  const VpzSetupHdr* hdr() const { return reinterpret_cast<const VpzSetupHdr*>(blob.data()); }
};

// Parses the identification header only (LoadStreamHeader).  0 or negative error.
int parse_id_header(const uint8_t* pkt, size_t len, IdHeader* out);

// Mode.GetPacketInfo given sizes and flags.
void compute_geometry(int size0, int size1, bool long_block, bool prev, bool next, PacketGeom* g);

// BlocksizeDerivedCache.CalcWindowSlope (BlocksizeDerivedCache.cs:24-35)
void window_slope(float* slope, int n);

uint64_t fnv1a64(const uint8_t* p, size_t n, uint64_t h = 1469598103934665603ull);

}  // namespace vpz
