"""Hand-written and fuzzed corpora for the parity tests (test infrastructure)."""
import random

import numpy as np

EXT = {"other": 0, "py": 1, "cc": 2, "cpp": 3, "java": 4, "c": 5, "h": 6}

PY_SAMPLE = b'''import unittest
from mock import patch

class SkillTest(object):
    """Assert that the gui can set gui variables."""
    def setUp(self):
        self.x = 1
        assert self.x

class TestThing(unittest.TestCase):
    def test_docker_agent_init(monkeypatch, runner_token):
        agent = DockerAgent()
        assert agent
        assert agent.labels == []
        assert agent.name == "agent"
        assert not agent.no_pull
        assert "Schedule not found" in str(exc.value)
        assert result == 0, "Repo did not pass Black formatting!"
        assert x is not None
        assert res.mapped == True
        assert loss[-1] < loss[0] * data.convergence
        assert sys.version_info >= (3, 6)
        assert a <= b
        assert a != b
        assert a > b
        assert(x)
        assert
        assert\t
    def test_more(self):
        self.assertEqual(a, b)
        self.assertEqual (a, b)
        self.assertEquals(a, b)
        self.assert_(a)
        self.assertListEqual(a, b)
        self.assertWeirdCustomThing(a)
        mock.assert_called_once_with(1)
        service.list_player.set_media_list.assert_called_with(x)
        parser.add_argument('-f', '--filename', dest='filename', default="/tmp/test.wav")
        except AssertionError:
        response.getTransform(assert_me)
        # TODO assert Service is Available
    async def test_set_multiple(self):
        x = GPUAssert(y)
classifier = 3
class\tTabbed:
'''

CC_SAMPLE = b'''#include "gtest/gtest.h"
#include "modules/perception/fusion/common/dst_evidence.h"

namespace apollo {
class DSTEvidenceTest : public ::testing::Test {
 public:
  DSTEvidenceTest()
      : sensor1_dst_("test"), sensor2_dst_("test"), fused_dst_("test") {
    dst_manager->AddApp("test", fod_subsets, fod_subset_names);
    vec_equal_ = [](const std::vector<double> &vec,
                    const std::vector<double> &gt) {
      CHECK_EQ(vec.size(), gt.size());
      for (size_t i = 0; i < vec.size(); ++i) {
        EXPECT_NEAR(vec[i], gt[i], 1e-6);
      }
    };
  }
  ~DSTEvidenceTest() {}
  void assign_dst_test() {
    EXPECT_DOUBLE_EQ(dst_vec[i], dst_vec_gt[i]);
  }
};
TEST_F(DSTEvidenceTest, assign_test) {
  ASSERT_TRUE(sensor1_dst_.SetBbaVec(sensor1_data));
  EXPECT_EQ(latest_observed_msg_ptr->class_name(), "BlockerTest");
  EXPECT_STREQ("a", "b");
  EXPECT_CALL(mock, Foo());
  EXPECT_THROW(f(), std::exception);
  EXPECT_FLOAT_EQ(1.0f, x);
  else ASSERT_EQ(1, 2);
  // EXPECT_EQ(a, b);
  RAPIDJSON_ASSERT(x);
  FOR_EACH(assertion, list) {
  static_assert(sizeof(int) == 4, "int");
  EXPECT_GE(a, b); EXPECT_LE(a, b);
}
  TEST_F(Indented, fixture) {
TEST(TestSuite, CheckGenerateAnchors) {
  void CreateTestMapNode(unsigned int m, unsigned int n,
  EXPECT_LT(a,
            b);
}
'''

JAVA_SAMPLE = b'''package org.deepspeech.libdeepspeech.test;
public class MapDecodeTest {
    @Test public void testDoubleInitialize() throws Exception {
        assertEquals("org.deepspeech.libdeepspeech.test", appContext.getPackageName());
        assert (audioFormat == 1); // 1 is PCM
        assertTrue(x);
        Assert.assertEquals(1, 2);
        assertNull(y);
        assertThat(z, is(1));
    }
    public void mobilityOperationEncodeTest() {
        assertArrayEquals(a, b);
    }
}
'''

EDGE_FILES = [
    (b"", 1), (b"\n", 1), (b"\n\n\n", 2), (b"a", 1), (b"a\n", 1), (b"a\nb", 2), (b"\r\n\r\n", 1),
    (b"assert x\r\nEXPECT_EQ(a, b);\r\n", 2), (b"assert", 1), (b"x" * 5000, 1),
    (b"x" * 4095 + b"\n", 2), (b"x" * 4096 + b"\n" + b"assert y\n", 1), (b"\n" * 5000, 2),
    (b"y" * 4090 + b"assert z == 1\nEXPECT_TRUE(q);\n", 1), (b"ASSERT_EQ(a,b);" * 1000, 2),
    (b"def test(self):\n" * 700, 1), (b"\x00\x01\xff\xfeassert\x80\n\x00", 1),
    (b"  \t  assert   x  ==  1   \t \n", 1), (b"self.assertEqual\n", 1), (b"EXPECT_\n", 2),
    (b"a" * 8200 + b" assert not x\n" + b"b" * 100 + b"\n", 1),
    (b"no newline at end assert x < 1", 1), (PY_SAMPLE, 1), (CC_SAMPLE, 2), (JAVA_SAMPLE, 4),
    (CC_SAMPLE, 3), (CC_SAMPLE, 5), (CC_SAMPLE, 6), (PY_SAMPLE, 0), (CC_SAMPLE, 1), (PY_SAMPLE, 2),
]

TOKENS = [b"assert", b"ASSERT_", b"Assert", b"assert ", b"EXPECT_", b"EXPECT_EQ", b"expect_", b"test", b"Test", b"TEST",
          b"TEST_F", b"TEST_F(", b"def", b"def ", b"class", b"class ", b"class\t", b"void", b"{", b"}", b"(", b")", b"self.",
          b"assertEqual", b"assertTrue", b"assert_", b"assert_called_with", b"assertFoo", b" not ", b" in ", b" is not ",
          b"True", b"==", b"!=", b"<=", b">=", b"<", b">", b"not ", b" ", b"  ", b"\t", b"\r", b"x", b"y", b"_", b".", b",",
          b"EQ", b"NE", b"NEAR", b"FLOAT_EQ", b"DOUBLE_EQ", b"THROW", b"STREQ", b"//", b"#", b'"', b"asser", b"ssert",
          b"EXPECT", b"tes", b"clas", b"voi", b"de", b"\x00", b"\xc3\xa9", b"0", b"9", b"assertassert", b"testtest",
          b"BOOST_CHECK", b"BOOST_CHECK_EQUAL", b"BOOST_CHECK(", b"NTA_CHECK(", b"TESTEQUAL", b"FAIL", b"_CHECK", b"_CHEC", b"TESTEQUA",
          b"!", b"(!", b"assert (", b"F", b"TEST_"]


def fuzz_file(rng: random.Random, size: int, nl_rate=0.08, long_lines=False) -> bytes:
    out = bytearray()
    while len(out) < size:
        r = rng.random()
        if r < nl_rate:
            out += b"\n"
        elif r < nl_rate + 0.02 and long_lines:
            out += bytes(rng.choice(b"abcdefgxyz ._(") for _ in range(rng.randrange(200, 6000)))
        else:
            out += rng.choice(TOKENS)
    out = bytes(out[:size])
    if rng.random() < 0.5 and out and not out.endswith(b"\n"):
        out = out[:-1] + b"\n"
    return out


def fuzz_corpus(seed: int, n_files: int, max_size: int, long_lines=False):
    rng = random.Random(seed)
    files, exts = [], []
    for i in range(n_files):
        kind = rng.random()
        if kind < 0.1:
            size = rng.randrange(0, 40)
        elif kind < 0.2:
            size = rng.choice([4095, 4096, 4097, 8191, 8192, 8193, 4096 + 239, 4096 + 240, 4096 + 241, 12288])
            size = min(size, max_size)
        else:
            size = rng.randrange(1, max_size)
        files.append(fuzz_file(rng, size, nl_rate=rng.choice([0.01, 0.05, 0.1, 0.3]), long_lines=long_lines))
        exts.append(rng.choice([0, 1, 1, 2, 2, 3, 4, 5, 6]))
    grps = [rng.randrange(0, 5) for _ in range(n_files)]
    return files, np.array(exts, np.uint8), np.array(grps, np.uint16)


def edge_corpus():
    files = [f for f, _ in EDGE_FILES]
    exts = np.array([e for _, e in EDGE_FILES], np.uint8)
    grps = np.array([i % 3 for i in range(len(files))], np.uint16)
    return files, exts, grps


def load_fixture(path):
    """Files of a tests/golden/*.npz corpus fixture (tools/make_golden.py: c1_fixture, c1_hazards)."""
    d = np.load(path)
    blob, size = d["blob"], d["size"].astype(np.int64)
    ends = np.cumsum(size)
    files = [blob[e - s:e].tobytes() for s, e in zip(size, ends)]
    grps = d["grp"].astype(np.uint16)
    return files, d["ext"].astype(np.uint8), grps, int(grps.max()) + 1 if len(grps) else 1


def score_g1(golden, names, files, events):
    """Golden G1 (ML-Testing-v1.xlsx!DeepSpeech rows of the bundled files): how many sheet statements the Rev-B events
    reproduce, and how much of their counts.  events: assertion events of a scan over `files`."""
    import collections
    by_file = collections.defaultdict(collections.Counter)
    for e in events:
        f = files[int(e["file"])]
        by_file[int(e["file"])][f[int(e["stmt_off"]):int(e["stmt_off"]) + int(e["stmt_len"])].decode("latin-1")] += 1
    idx = {n: i for i, n in enumerate(names)}
    stm, cnt, per_file = [0, 0], [0, 0], {}
    for name, want in golden.items():
        got = by_file[idx[name]]
        fs = fc = 0
        for st, (c, _) in want.items():
            stm[1] += 1
            cnt[1] += c
            if got.get(st):
                stm[0] += 1
                cnt[0] += min(c, got[st])
                fs += 1
                fc += min(c, got[st])
        per_file[name] = ([fs, len(want)], [fc, sum(c for c, _ in want.values())])
    return stm, cnt, per_file


def load_fixture_names(path):
    d = np.load(path)
    return bytes(d["names"]).decode().split("\n")
