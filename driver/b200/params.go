package b200

import "math/big"

func mustHex(s string) *big.Int {
	v, ok := new(big.Int).SetString(s, 16)
	if !ok {
		panic("bad constant")
	}
	return v
}

func mustBytes(size int, vals ...string) []byte {
	out := make([]byte, 0, size*len(vals))
	for _, s := range vals {
		b := make([]byte, size)
		mustHex(s).FillBytes(b)
		out = append(out, b...)
	}
	return out
}

// group orders: reference math_test.go:261-270
var (
	orderBN254    = mustHex("30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001")
	orderBLS12381 = mustHex("73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001")
	orderBLS12377 = mustHex("12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001")
)

var (
	modBN254    = mustHex("30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47")
	modBLS12381 = mustHex("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab")
	modBLS12377 = mustHex("01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001")
)

func fieldModulus(id int) *big.Int {
	switch id {
	case idBN254:
		return modBN254
	case idBLS12377Gurvy:
		return modBLS12377
	}
	return modBLS12381
}

// generators in the Bytes() encoding; G1 from reference math_test.go:250-259, G2 the standard ones (SURVEY A.1);
// G2 wire order is X.A1 || X.A0 || Y.A1 || Y.A0.
var (
	g1GenBN254 = mustBytes(32, "1", "2")
	g2GenBN254 = mustBytes(32,
		"198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2",
		"1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed",
		"090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b",
		"12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa")
	g1GenBLS12381 = mustBytes(48,
		"17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb",
		"08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1")
	g2GenBLS12381 = mustBytes(48,
		"13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e",
		"024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8",
		"0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be",
		"0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801")
	g1GenBLS12377 = mustBytes(48,
		"008848defe740a67c8fc6225bf87ff5485951e2caa9d41bb188282c8bd37cb5cd5481512ffcd394eeab9b16eb21be9ef",
		"01914a69c5102eff1f674f5d30afeec4bd7fb348ca3e52d96d182ad44fb82305c2fe3d3634a9591afd82de55559c8ea6")
	g2GenBLS12377 = mustBytes(48,
		"00ea6040e700403170dc5a51b1b140d5532777ee6651cecbe7223ece0799c9de5cf89984bff76fe6b26bfefa6ea16afe",
		"018480be71c785fec89630a2a3841d01c565f071203e50317ea501f557db6b9b71889f52bb53540274e3e48f7c005196",
		"00f8169fd28355189e549da3151a70aa61ef11ac3d591bf12463b01acee304c24279b83f5e52270bd9a1cdd185eb8f93",
		"00690d665d446f7bd960736bcbb2efb4de03ed7274b49a58e458c282f832d204f2cf88886d8c7c2ef094094409fd4ddf")
)
