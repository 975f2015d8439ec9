// Row f3 of the scope table: PDB text -> the padded per-structure arrays the hot path consumes.
//
// Host-only code (no kernels).  Replaces what the reference does through biotite + pandas in
// PDB.read_pdb / tidy_structure / PDB._initialize_lookup / PDB._compute_atom_xyz
// (protstruc/pdb.py:24-40, 55-151): first MODEL only, first alternate location per residue, non-standard
// residue names substituted, only the 20 canonical amino acids and standard heavy-atom names kept,
// residues in file order with UNK placeholders for numbering gaps inside a chain, chain index by first
// appearance, atom slot = position of the atom name in the residue type's 15-slot list.
// One pass over the text, no allocation beyond small fixed tables; thread-safe (no global state), so the
// Python side parses many files concurrently (ctypes releases the GIL).

#include <cstdlib>
#include <cstring>
#include <cmath>

#include "common.cuh"

namespace ps {

namespace {

constexpr int kSlots = 15;

struct ResidueType {
    const char* name;
    char one;
    const char* atoms[kSlots];
};

// Heavy-atom slot table (the AlphaFold-style atom14 + OXT layout the reference uses, general.py:149-171).
const ResidueType kTypes[20] = {
    {"ALA", 'A', {"N", "CA", "C", "O", "CB", "", "", "", "", "", "", "", "", "", "OXT"}},
    {"ARG", 'R', {"N", "CA", "C", "O", "CB", "CG", "CD", "NE", "CZ", "NH1", "NH2", "", "", "", "OXT"}},
    {"ASN", 'N', {"N", "CA", "C", "O", "CB", "CG", "OD1", "ND2", "", "", "", "", "", "", "OXT"}},
    {"ASP", 'D', {"N", "CA", "C", "O", "CB", "CG", "OD1", "OD2", "", "", "", "", "", "", "OXT"}},
    {"CYS", 'C', {"N", "CA", "C", "O", "CB", "SG", "", "", "", "", "", "", "", "", "OXT"}},
    {"GLN", 'Q', {"N", "CA", "C", "O", "CB", "CG", "CD", "OE1", "NE2", "", "", "", "", "", "OXT"}},
    {"GLU", 'E', {"N", "CA", "C", "O", "CB", "CG", "CD", "OE1", "OE2", "", "", "", "", "", "OXT"}},
    {"GLY", 'G', {"N", "CA", "C", "O", "", "", "", "", "", "", "", "", "", "", "OXT"}},
    {"HIS", 'H', {"N", "CA", "C", "O", "CB", "CG", "ND1", "CD2", "CE1", "NE2", "", "", "", "", "OXT"}},
    {"ILE", 'I', {"N", "CA", "C", "O", "CB", "CG1", "CG2", "CD1", "", "", "", "", "", "", "OXT"}},
    {"LEU", 'L', {"N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "", "", "", "", "", "", "OXT"}},
    {"LYS", 'K', {"N", "CA", "C", "O", "CB", "CG", "CD", "CE", "NZ", "", "", "", "", "", "OXT"}},
    {"MET", 'M', {"N", "CA", "C", "O", "CB", "CG", "SD", "CE", "", "", "", "", "", "", "OXT"}},
    {"PHE", 'F', {"N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "CE1", "CE2", "CZ", "", "", "", "OXT"}},
    {"PRO", 'P', {"N", "CA", "C", "O", "CB", "CG", "CD", "", "", "", "", "", "", "", "OXT"}},
    {"SER", 'S', {"N", "CA", "C", "O", "CB", "OG", "", "", "", "", "", "", "", "", "OXT"}},
    {"THR", 'T', {"N", "CA", "C", "O", "CB", "OG1", "CG2", "", "", "", "", "", "", "", "OXT"}},
    {"TRP", 'W', {"N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "NE1", "CE2", "CE3", "CZ2", "CZ3", "CH2", "OXT"}},
    {"TYR", 'Y', {"N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "CE1", "CE2", "CZ", "OH", "", "", "OXT"}},
    {"VAL", 'V', {"N", "CA", "C", "O", "CB", "CG1", "CG2", "", "", "", "", "", "", "", "OXT"}},
};

// Non-standard -> standard residue names (the OpenMM/PDBFixer table the reference embeds,
// general.py:109-124), grouped by target.
struct Substitution {
    const char* target;
    const char* sources;  // space separated
};
const Substitution kSubstitutions[] = {
    {"ALA", "AIB ALM AYA BNN CHG CSD DAL DHA DNP FLA HAC MAA PRR TIH TPQ"},
    {"ARG", "ACL AGM ARM DAR HAR HMR"},
    {"ASN", "MEN"},
    {"ASP", "2AS ASA ASB ASK ASL ASQ BHD DAS DSP IAS"},
    {"CYS", "BCS BUC C5C C6C CAS CCS CEA CME CSO CSP CSS CSW CSX CY1 CY3 CYG CYM CYQ DCY EFC OCS PEC PR3 PYX SCH SCS SCY SHC SMC SOC"},
    {"GLN", "DGN"},
    {"GLU", "5HP CGU DGL GGL GMA PCA"},
    {"GLY", "GL3 GLZ GSC MPQ MSA NMC SAR"},
    {"HIS", "3AH DHI HIC HIP MHS NEM NEP"},
    {"ILE", "DIL IIL"},
    {"LEU", "BUG CLE DLE MLE NLE NLN NLP"},
    {"LYS", "ALY DLY KCX LLP LLY LYM LYZ SHR TRG"},
    {"MET", "CXM FME MSE OMT"},
    {"PHE", "DAH DPN HPQ PHI PHL"},
    {"PRO", "DPR HYP"},
    {"SER", "DSN MIS OAS SAC SEL SEP SET SVA"},
    {"THR", "ALO BMT DTH TPO"},
    {"TRP", "DTR HTR LTR TPL TRO"},
    {"TYR", "DTY IYR PAQ PTR STY TYB TYI TYQ TYS TYY"},
    {"VAL", "DIV DVA MVA"},
};

// Copies columns [begin, end) of a line without surrounding blanks into out (NUL terminated).
void field(const char* line, int line_len, int begin, int end, char* out, int out_cap) {
    if (end > line_len) end = line_len;
    while (begin < end && line[begin] == ' ') ++begin;
    while (end > begin && (line[end - 1] == ' ' || line[end - 1] == '\r')) --end;
    int n = end - begin;
    if (n < 0) n = 0;
    if (n > out_cap - 1) n = out_cap - 1;
    memcpy(out, line + begin, n);
    out[n] = '\0';
}

// Names are at most four characters: they are compared as packed 32-bit keys (first character in the low byte).
uint32_t pack_name(const char* s) {
    uint32_t key = 0;
    for (int k = 0; k < 4 && s[k]; ++k) key |= static_cast<uint32_t>(static_cast<unsigned char>(s[k])) << (8 * k);
    return key;
}
// Key of the trimmed text in columns [begin, end) (end - begin <= 4); 0 when blank.  Returns false when the
// trimmed text has an inner blank (never a valid name).
bool pack_columns(const char* line, int begin, int end, uint32_t* key) {
    while (begin < end && line[begin] == ' ') ++begin;
    while (end > begin && (line[end - 1] == ' ' || line[end - 1] == '\r')) --end;
    uint32_t k = 0;
    for (int c = begin; c < end; ++c) {
        if (line[c] == ' ') return false;
        k |= static_cast<uint32_t>(static_cast<unsigned char>(line[c])) << (8 * (c - begin));
    }
    *key = k;
    return true;
}

// Lookup tables built once (thread-safe function-local static): residue name -> type (standard names and the
// substitution table folded in), the set of heavy-atom names, and the packed slot names of every type.
struct Tables {
    static constexpr int kResBits = 10, kAtomBits = 8;
    uint32_t res_key[1 << kResBits];
    int8_t res_type[1 << kResBits];
    uint32_t atom_key[1 << kAtomBits];
    uint32_t slot_key[20][kSlots];

    static uint32_t hash(uint32_t key, int bits) { return (key * 2654435761u) >> (32 - bits); }

    void add_residue(uint32_t key, int type) {
        uint32_t h = hash(key, kResBits);
        while (res_key[h] != 0 && res_key[h] != key) h = (h + 1) & ((1u << kResBits) - 1);
        if (res_key[h] == 0) {  // first entry wins, like the linear scans this replaces
            res_key[h] = key;
            res_type[h] = static_cast<int8_t>(type);
        }
    }
    Tables() {
        memset(res_key, 0, sizeof res_key);
        memset(res_type, -1, sizeof res_type);
        memset(atom_key, 0, sizeof atom_key);
        for (int t = 0; t < 20; ++t) add_residue(pack_name(kTypes[t].name), t);
        for (const Substitution& sub : kSubstitutions) {
            int target = -1;
            for (int t = 0; t < 20; ++t)
                if (strcmp(kTypes[t].name, sub.target) == 0) target = t;
            const char* p = sub.sources;
            while (*p) {
                char name[4] = {p[0], p[1], p[2], '\0'};
                add_residue(pack_name(name), target);
                p += 3;
                while (*p == ' ') ++p;
            }
        }
        for (int t = 0; t < 20; ++t)
            for (int a = 0; a < kSlots; ++a) {
                const uint32_t key = pack_name(kTypes[t].atoms[a]);
                slot_key[t][a] = key;
                if (key == 0) continue;
                uint32_t h = hash(key, kAtomBits);
                while (atom_key[h] != 0 && atom_key[h] != key) h = (h + 1) & ((1u << kAtomBits) - 1);
                atom_key[h] = key;
            }
    }
    // Type of a residue name after substitution (-1: not one of the 20 amino acids).
    int residue_type(uint32_t key) const {
        if (key == 0) return -1;
        uint32_t h = hash(key, kResBits);
        while (res_key[h] != 0) {
            if (res_key[h] == key) return res_type[h];
            h = (h + 1) & ((1u << kResBits) - 1);
        }
        return -1;
    }
    bool is_heavy_atom(uint32_t key) const {
        if (key == 0) return false;
        uint32_t h = hash(key, kAtomBits);
        while (atom_key[h] != 0) {
            if (atom_key[h] == key) return true;
            h = (h + 1) & ((1u << kAtomBits) - 1);
        }
        return false;
    }
    int slot_of(int type, uint32_t key) const {
        for (int a = 0; a < kSlots; ++a)
            if (slot_key[type][a] == key) return a;
        return -1;
    }
};

const Tables& tables() {
    static const Tables t;
    return t;
}

// Integer in columns [begin, end) with atoi semantics (leading blanks, optional sign, digits until anything else).
int parse_int(const char* line, int begin, int end) {
    while (begin < end && line[begin] == ' ') ++begin;
    bool negative = false;
    if (begin < end && (line[begin] == '-' || line[begin] == '+')) negative = line[begin++] == '-';
    int v = 0;
    while (begin < end && line[begin] >= '0' && line[begin] <= '9') v = v * 10 + (line[begin++] - '0');
    return negative ? -v : v;
}

// Fixed-point coordinate field ("%8.3f").  digits / 10^decimals with both operands exact in double is the
// correctly rounded double of the decimal text, i.e. what strtod returns; anything that is not plain
// [sign]digits[.digits] goes to strtod itself.
float parse_coordinate(const char* line, int line_len, int begin, int end) {
    static const double kPow10[10] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9};
    if (end > line_len) end = line_len;
    int c = begin;
    while (c < end && line[c] == ' ') ++c;
    int stop = end;
    while (stop > c && (line[stop - 1] == ' ' || line[stop - 1] == '\r')) --stop;
    int q = c;
    bool negative = false;
    if (q < stop && (line[q] == '-' || line[q] == '+')) negative = line[q++] == '-';
    long long mantissa = 0;
    int digits = 0, decimals = 0;
    while (q < stop && line[q] >= '0' && line[q] <= '9') {
        mantissa = mantissa * 10 + (line[q++] - '0');
        ++digits;
    }
    if (q < stop && line[q] == '.') {
        ++q;
        while (q < stop && line[q] >= '0' && line[q] <= '9') {
            mantissa = mantissa * 10 + (line[q++] - '0');
            ++digits;
            ++decimals;
        }
    }
    if (q == stop && digits > 0 && digits <= 15 && decimals <= 9) {
        const double v = static_cast<double>(mantissa) / kPow10[decimals];
        return static_cast<float>(negative ? -v : v);
    }
    char text[16];
    field(line, line_len, begin, end, text, sizeof text);
    return static_cast<float>(strtod(text, nullptr));
}

struct Sink {
    int capacity;       // rows available in the output arrays (0 = counting only)
    float* xyz;
    uint8_t* mask;
    int32_t* chain_idx;
    char* chain_id;
    int32_t* resseq;
    char* icode;
    char* aa1;
    int rows = 0;
    char chains[64];
    int n_chains = 0;

    int chain_index(char c) {
        for (int k = 0; k < n_chains; ++k)
            if (chains[k] == c) return k;
        if (n_chains < 64) chains[n_chains] = c;
        return n_chains++;
    }
    // Opens a new row; returns its index (or -1 when only counting / out of capacity).
    int open_row(char chain, int number, char ins, char one) {
        const int idx = rows++;
        const int ci = chain_index(chain);
        if (idx >= capacity) return -1;
        const float nan = nanf("");
        for (int k = 0; k < kSlots * 3; ++k) xyz[idx * kSlots * 3 + k] = nan;
        memset(mask + idx * kSlots, 0, kSlots);
        chain_idx[idx] = ci;
        chain_id[idx] = chain;
        resseq[idx] = number;
        icode[idx] = ins;
        aa1[idx] = one;
        return idx;
    }
};

int parse(const char* text, long long len, Sink& sink) {
    const Tables& tb = tables();
    bool seen_model = false, in_first_model = true;
    // identity of the residue currently being filled
    bool have_residue = false;
    char cur_chain = 0, cur_ins = 0;
    int cur_number = 0, cur_type = -1, cur_row = -1;
    char first_altloc = 0;
    // gap bookkeeping (reference pdb.py:92-121)
    bool have_chain = false;
    char gap_chain = 0;
    int gap_number = 0;

    long long pos = 0;
    while (pos < len) {
        const char* line = text + pos;
        const char* nl = static_cast<const char*>(memchr(line, '\n', len - pos));
        const int line_len = static_cast<int>(nl ? nl - line : len - pos);
        pos += line_len + 1;
        if (line_len < 6) continue;
        if (strncmp(line, "MODEL", 5) == 0) {
            if (seen_model) in_first_model = false;
            seen_model = true;
            continue;
        }
        if (strncmp(line, "ENDMDL", 6) == 0) {
            in_first_model = false;
            continue;
        }
        if (!in_first_model) continue;
        if (strncmp(line, "ATOM  ", 6) != 0 && strncmp(line, "HETATM", 6) != 0) continue;
        if (line_len < 54) continue;

        uint32_t atom_key = 0, res_key = 0;
        if (!pack_columns(line, 12, 16, &atom_key) || !pack_columns(line, 17, 20, &res_key)) continue;
        const int type = tb.residue_type(res_key);
        if (type < 0 || !tb.is_heavy_atom(atom_key)) continue;
        const char altloc = line[16] == ' ' ? 0 : line[16];
        const char chain = line[21];
        const char ins = line[26] == ' ' ? 0 : line[26];
        const int number = parse_int(line, 22, 26);

        // (after substitution a residue name is one of the 20 standard ones, so comparing types compares names)
        const bool same = have_residue && chain == cur_chain && number == cur_number && ins == cur_ins &&
                          type == cur_type;
        if (!same) {
            // a new residue starts: fill numbering gaps inside the chain with UNK rows first
            if (!have_chain || gap_chain != chain) {
                gap_chain = chain;
                gap_number = number;
                have_chain = true;
            }
            while (gap_number + 1 < number) {
                sink.open_row(gap_chain, gap_number + 1, ins, 'X');
                ++gap_number;
            }
            cur_row = sink.open_row(chain, number, ins, kTypes[type].one);
            gap_chain = chain;
            gap_number = number;
            have_residue = true;
            cur_chain = chain;
            cur_number = number;
            cur_ins = ins;
            cur_type = type;
            first_altloc = 0;
        }
        if (altloc) {  // keep only the first alternate location seen in this residue
            if (!first_altloc) first_altloc = altloc;
            if (altloc != first_altloc) continue;
        }
        const int slot = tb.slot_of(cur_type, atom_key);
        if (slot < 0 || cur_row < 0) continue;
        float* dst = sink.xyz + (static_cast<long long>(cur_row) * kSlots + slot) * 3;
        for (int k = 0; k < 3; ++k) dst[k] = parse_coordinate(line, line_len, 30 + 8 * k, 38 + 8 * k);
        sink.mask[cur_row * kSlots + slot] = 1;
    }
    return sink.rows;
}

}  // namespace

int host_pdb_parse_impl(const char* text, long long len, int capacity, float* xyz, uint8_t* mask,
                        int32_t* chain_idx, char* chain_id, int32_t* resseq, char* icode, char* aa1,
                        int* n_residues) {
    PS_REQUIRE(text != nullptr && len >= 0, PS_ERR_NULL_POINTER, "pdb_parse: text is NULL");
    PS_REQUIRE(n_residues != nullptr, PS_ERR_NULL_POINTER, "pdb_parse: n_residues is NULL");
    PS_REQUIRE(capacity == 0 || (xyz && mask && chain_idx && chain_id && resseq && icode && aa1),
               PS_ERR_NULL_POINTER, "pdb_parse: output arrays are required when capacity > 0");
    Sink sink{};
    sink.capacity = capacity;
    sink.xyz = xyz;
    sink.mask = mask;
    sink.chain_idx = chain_idx;
    sink.chain_id = chain_id;
    sink.resseq = resseq;
    sink.icode = icode;
    sink.aa1 = aa1;
    *n_residues = parse(text, len, sink);
    PS_REQUIRE(capacity == 0 || *n_residues <= capacity, PS_ERR_BAD_SHAPE,
               "pdb_parse: %d residues do not fit in the %d rows provided", *n_residues, capacity);
    return PS_OK;
}

}  // namespace ps
