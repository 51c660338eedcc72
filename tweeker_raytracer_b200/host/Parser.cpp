#include "Parser.h"

#include <cctype>
#include <fstream>
#include <iostream>
#include <sstream>

bool Parser::load(std::string const& filename)
{
  m_source.clear(); m_index = 0; m_line = 1;
  std::ifstream in(filename);
  if (!in) { std::cerr << "ERROR: loadString() Failed to open file " << filename << std::endl; return false; }
  std::stringstream data;
  data << in.rdbuf();
  if (in.fail()) { std::cerr << "ERROR: loadString() Failed to read file " << filename << std::endl; return false; }
  m_source = data.str();
  return true;
}

static inline bool is_blank(char c) { return c == ' ' || c == '\t'; }
static inline bool is_delim(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; }

ParserTokenType Parser::getNextToken(std::string& token)
{
  token.clear();
  const std::string::size_type n = m_source.size();
  for (;;)
  {
    while (m_index < n && is_blank(m_source[m_index])) ++m_index;
    if (m_index >= n) return PTT_EOF;
    const char c = m_source[m_index];
    if (c == '#')
    {
      while (m_index < n && m_source[m_index] != '\n') ++m_index;
      if (m_index >= n) return PTT_EOF;
      ++m_index; ++m_line;
      continue;
    }
    if (c == '\r') { ++m_index; continue; }
    if (c == '\n') { ++m_index; ++m_line; continue; }
    std::string::size_type last = m_index;
    while (last < n && !is_delim(m_source[last])) ++last;
    token = m_source.substr(m_index, last - m_index);
    m_index = last;
    if (std::isdigit((unsigned char)c) || c == '-' || c == '+' || c == '.')
    {
      if (token.find_first_not_of("+-0123456789.eE") == std::string::npos) return PTT_VAL;
    }
    return PTT_ID;
  }
}

ParserTokenType Parser::getNextLine(std::string& token)
{
  token.clear();
  const std::string::size_type n = m_source.size();
  while (m_index < n && is_blank(m_source[m_index])) ++m_index;
  if (m_index >= n) return PTT_EOF;
  const char c = m_source[m_index];
  if (c == '\r') { ++m_index; return PTT_EOL; }
  if (c == '\n') { ++m_index; ++m_line; return PTT_EOL; }
  std::string::size_type last = m_index;
  while (last < n && m_source[last] != '\r' && m_source[last] != '\n') ++last;
  const std::string::size_type first = m_index;
  m_index = last;
  while (first < last && is_delim(m_source[last - 1])) --last;
  if (first == last) return PTT_EOL;
  token = m_source.substr(first, last - first);
  return PTT_ID;
}
